#!/bin/bash
# One B200: every bench line, tool bench and ncu capture that profiles/ and DESIGN.md quote.  Outputs: gpurun_out/m_*
cd "$(dirname "$0")/.."
O=gpurun_out
T=${1:-m}
python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -1 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
for w in c2 c1 c4 n1m c3; do
  python bench.py --workload $w --steps 30 --warmup 5 > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err
  python -c "import json; d=json.load(open('$O/${T}_bench_$w.json')); print('$w', '%.4g' % d['value'], d['ms_per_step'], d['roofline']['frac'] if d.get('roofline') else None, d['e2e']['value'] if d.get('e2e') else None, d['e2e'].get('route') if d.get('e2e') else None)"
done
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref_c2.json 2> $O/${T}_bench_ref_c2.err; head -c 300 $O/${T}_bench_ref_c2.json; echo
python tools/bench_kernels.py > $O/${T}_kernels.json 2> $O/${T}_kernels.err
python tools/bench_skyh5.py > $O/${T}_skyh5.json 2> $O/${T}_skyh5.err
python tools/bench_sweep.py > $O/${T}_sweep.json 2> $O/${T}_sweep.err; cat $O/${T}_sweep.json
python tools/bench_sweep.py --lmn > $O/${T}_sweep_lmn.json 2>> $O/${T}_sweep.err
# ncu: launch list of the bench command, then one full capture per dominant kernel (only after the plain runs above)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval -c 2 -o $O/${T}_prof_k_eval_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > $O/${T}_ncu_keval.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_positions_sweep -c 1 -o $O/${T}_prof_sweep python tools/bench_sweep.py --t 256 --reps 1 --cpu-sample 100 > $O/${T}_ncu_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval -c 1 -o $O/${T}_prof_apparent python tools/prof_apparent.py > $O/${T}_ncu_apparent.log 2>&1
ls -la $O | grep ${T}_ | awk '{print $5, $9}'
