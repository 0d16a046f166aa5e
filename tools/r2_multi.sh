#!/bin/bash
# multi-GPU check: gpurun --gpus N --timeout 1500 -- 'bash tools/r2_multi.sh <tag> N [parity] [c5] [c4] [ref]'
export TAG=${1:-r2_multi}; N=${2:-2}; shift; shift
STAGES=${*:-parity c5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if has parity; then
  timeout 600 $TR --master-port 29511 tools/sharded_parity.py > $OUT/parity_peer_n$N.txt 2>&1; echo "parity peer rc=$?"; grep -E "OK|MISMATCH" $OUT/parity_peer_n$N.txt
  AGP_NO_PEER=1 timeout 600 $TR --master-port 29512 tools/sharded_parity.py > $OUT/parity_nccl_n$N.txt 2>&1; echo "parity nccl rc=$?"; grep -E "OK|MISMATCH" $OUT/parity_nccl_n$N.txt
fi
if has c5; then
  timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 > $OUT/bench_c5_n$N.json 2> $OUT/bench_c5_n$N.err; echo "c5 rc=$?"; tail -c 400 $OUT/bench_c5_n$N.err
fi
if has c4; then
  timeout 600 $TR --master-port 29514 bench.py --gpus $N --config C4 --steps 50 --warmup 3 > $OUT/bench_c4_n$N.json 2> $OUT/bench_c4_n$N.err; echo "c4 rc=$?"; tail -c 400 $OUT/bench_c4_n$N.err
fi
if has ref; then
  timeout 600 python bench.py --impl reference --gpus $N --steps 4 --warmup 1 > $OUT/ref_c5.json 2> $OUT/ref_c5.err; echo "ref rc=$?"
fi
python - <<'PY'
import glob, json, os
for f in sorted(glob.glob(os.path.join("gpurun_out", os.environ["TAG"], "*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), "N", d["n_gpus"], round(d["value"]), "latent-it/s;", round(d["ms_per_step"] * 1e3, 1), "us/step; e2e", (d.get("e2e") or {}).get("value"), (d.get("run") or {}).get("exchange"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
