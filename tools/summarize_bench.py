#!/usr/bin/env python
"""Markdown tables from bench.py JSON lines (profiles/r02_bench_*gpu.json): one row per config and GPU count."""
import json
import sys


def load(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def main(paths):
    runs = sorted((load(p) for p in paths), key=lambda d: d["n_gpus"])
    ns = [d["n_gpus"] for d in runs]
    print("| workload | " + " | ".join(f"{n} GPU" + ("s" if n > 1 else "") for n in ns) + " | efficiency at %d |" % ns[-1])
    print("|---|" + "---|" * (len(ns) + 1))

    def row(name, get):
        vals = [get(d) for d in runs]
        cells = []
        for v in vals:
            cells.append("—" if v is None else f"{v[0]:.3g} ({v[1]:.3f} ms, frac {v[2]:.2f})")
        eff = "—"
        if vals[0] and vals[-1] and ns[0] == 1:
            eff = f"{vals[-1][0] / vals[0][0] / ns[-1]:.2f}"
        print(f"| {name} | " + " | ".join(cells) + f" | {eff} |")

    row("n1m headline (evals/s)", lambda d: (d["value"], d["ms_per_step"], d["roofline"]["frac"]))
    for key in ("c2", "c3", "c4", "full", "c5"):
        def get(d, key=key):
            c = (d.get("configs") or {}).get(key)
            if not c or "error" in c:
                return None
            return (c["value"], c["ms_per_step"], c["roofline"]["frac"])
        row(key + (" (source·times/s)" if key == "c5" else " (evals/s)"), get)
    print()
    print("| GPUs | e2e evals/s | ms/step | route | D2H all ranks GB/s | expand all ranks GB/s out | frac of binding floor | visibility e2e evals/s |")
    print("|---|---|---|---|---|---|---|---|")
    for d in runs:
        e = d.get("e2e")
        if not e:
            continue
        r = e["roofline"]
        v = e.get("visibility") or {}
        print(f"| {d['n_gpus']} | {e['value']:.3g} | {e['ms_per_step']:.0f} | {e['route']} | {r['d2h_GBps_all_ranks']:.0f} | "
              f"{r['expand_GBps_out_all_ranks']:.0f} | {r['frac_of_binding_floor']:.2f} | {v.get('value', float('nan')):.3g} |")
    print()
    print("| GPUs | gathered GB | peer copy rate GB/s | copy route all-gather ms (GB/s recv, frac) | copy route to rank 0 | store route all-gather | NCCL all_gather + re-pack |")
    print("|---|---|---|---|---|---|---|")
    for d in runs:
        g = d.get("gather")
        if not g or "error" in g:
            continue

        def cell(k, key="GBps_received_per_gpu"):
            x = g.get(k)
            if not x:
                return "—"
            return f"{x['ms']:.1f} ({x.get(key, x.get('GBps_received')):.0f}, {x['frac_of_peer_copy']:.2f})"
        print(f"| {d['n_gpus']} | {g['bytes_gathered'] / 1e9:.2f} | {g['peer_copy_GBps_per_gpu']:.0f} | {cell('peer_copy_all_gather')} | "
              f"{cell('peer_copy_to_rank0', 'GBps_received')} | {cell('peer_store_all_gather')} | {cell('nccl_all_gather_and_repack')} |")


if __name__ == "__main__":
    main(sys.argv[1:])
