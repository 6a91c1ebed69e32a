python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02_items_tests.log; cat gpurun_out/r02_items_tests.log
python tools/diag_timeline.py 8 2>&1 | tail -3 > gpurun_out/r02_timeline_v6.txt; cat gpurun_out/r02_timeline_v6.txt
python bench.py --steps 20 --warmup 5 --no-extra 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'pair us', round(d['roofline']['us_per_launch'],1), 'frac', round(d['roofline']['frac'],4))"
