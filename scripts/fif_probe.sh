#!/bin/bash
# throughput vs footprint knobs of the front kernels
run() { echo "== $*"; env "$@" timeout 240 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'p50', d['frame_ms_p50'], 'e2e', d['e2e']['value'], [s['us'] for s in d['stages']])"; }
run BGS_FRAMES_IN_FLIGHT=3
run BGS_FRAMES_IN_FLIGHT=3 BGS_COOP_BLOCKS=1
run BGS_FRAMES_IN_FLIGHT=3 BGS_COOP_BLOCKS=2
run BGS_FRAMES_IN_FLIGHT=3 BGS_COOP_BLOCKS=3
run BGS_FRAMES_IN_FLIGHT=3 BGS_COOP_BLOCKS=2 BGS_PROJECT_CTAS=2
run BGS_FRAMES_IN_FLIGHT=3 BGS_COOP_BLOCKS=2 BGS_PROJECT_CTAS=3
run BGS_FRAMES_IN_FLIGHT=4 BGS_COOP_BLOCKS=2 BGS_PROJECT_CTAS=3
run BGS_FRAMES_IN_FLIGHT=3 BGS_COOP_BLOCKS=2 BGS_RASTER_PRIO=0
