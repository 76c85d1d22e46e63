#!/bin/bash
run() { echo "== $*"; env "$@" timeout 240 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'p50', d['frame_ms_p50'], 'e2e', d['e2e']['value'], [s['us'] for s in d['stages']])"; }
run BGS_SORT_CTAS_ASYNC=1
run BGS_SORT_CTAS_ASYNC=2
run BGS_SORT_CTAS_ASYNC=1 BGS_COOP_BLOCKS_ASYNC=1
run BGS_SORT_CTAS_ASYNC=1 BGS_PROJECT_CTAS=4
