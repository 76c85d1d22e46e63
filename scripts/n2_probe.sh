#!/bin/bash
run() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 6 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'ce', d['gather_ce']['value'], 'ok', d['gathered_frames_verified'])"; }
run BGS_FRAMES_IN_FLIGHT=4
run BGS_FRAMES_IN_FLIGHT=6
