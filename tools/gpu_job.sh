(RPB_DEBUG_GRAPH=1 python tools/diag_hop_graph.py 150) > gpurun_out/r02_hop_graph3.log 2>&1; cat gpurun_out/r02_hop_graph3.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-extra 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'pair us', round(d['roofline']['us_per_launch'],1), 'frac', round(d['roofline']['frac'],4))"
