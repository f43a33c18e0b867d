#!/bin/bash
for rep in 1 2; do
python scripts_dev/bench_r1.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('old bench: ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
python bench.py --no-cpu-baseline --lean --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('new bench: ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done
