#!/bin/bash
# First GPU session of the next round: everything that was written after the round-1 GPU budget ran out,
# cheapest first, each step under its own timeout so a hang costs one step, not the box.
#   gpurun --timeout 1500 -- 'bash tools/next_gpu_session.sh'            (1 GPU: steps 1-3)
#   gpurun --gpus 4 --timeout 900 -- 'bash tools/next_gpu_session.sh multi'   (step 4)
set -x
mkdir -p gpurun_out
if [ "$1" != "multi" ]; then
  # 0. building-block rates the next kernel designs hinge on: seconds of GPU time
  for part in single pair; do
    for grid in 1 148; do
      timeout 120 python tools/ubench.py --part $part --grid $grid >> gpurun_out/ubench.txt 2>&1
    done
  done
  cat gpurun_out/ubench.txt
  timeout 120 python tools/probe_pair.py > gpurun_out/probe_pair.txt 2>&1; cat gpurun_out/probe_pair.txt
  # 1. full parity suite incl. the head_dim-64 (C1) tests that have not run on a GPU yet
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
  # 2. the CTA-pair forward variants under the forward + API tests (gate for switching the default)
  for impl in 5 6; do
    BA_FWD_IMPL=$impl timeout 600 python -m pytest tests/test_gpu_fwd.py tests/test_gpu_api.py -x -q \
      > gpurun_out/pytest_fwd_impl$impl.txt 2>&1; tail -2 gpurun_out/pytest_fwd_impl$impl.txt
  done
  # 3. ncu --set full of variants 2 (default), 5, 6 at one launch size: where do the pair kernels stall?
  for impl in 2 5 6; do
    BA_FWD_IMPL=$impl timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd -s 2 -c 1 \
      -o gpurun_out/prof_fwd_impl$impl -f python bench.py --seq 16384 --steps 1 --warmup 3 --no-e2e --no-cpu \
      > gpurun_out/prof_fwd_impl$impl.out 2>&1
  done
else
  # 4. hierarchical (double) ring over NCCL: parity first, then exposed time next to the flat ring
  N=$(nvidia-smi -L | wc -l)
  RING_CHECK_DOUBLE=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N \
    --master-addr 127.0.0.1 --master-port 29541 tests/ring_check.py > gpurun_out/ring_check_double_n$N.txt 2>&1
  tail -14 gpurun_out/ring_check_double_n$N.txt
  for dr in 0 2; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
      --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --double-ring $dr \
      > gpurun_out/bench_n${N}_double$dr.json 2> gpurun_out/bench_n${N}_double$dr.err
    tail -1 gpurun_out/bench_n${N}_double$dr.json
  done
  # 5. copy-engine transport (never run before): parity at N ranks under a SHORT timeout, then the short-shard
  #    bench point where NCCL's hop is exposed (S_local = 65536/N), both transports
  BA_RING_TRANSPORT=ce timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N \
    --master-addr 127.0.0.1 --master-port 29543 tests/ring_check.py > gpurun_out/ring_check_ce_n$N.txt 2>&1
  tail -8 gpurun_out/ring_check_ce_n$N.txt
  if grep -q FAIL gpurun_out/ring_check_ce_n$N.txt || ! grep -q PASS gpurun_out/ring_check_ce_n$N.txt; then
    echo "copy-engine ring not healthy: skipping its bench"
  else
    for tr in nccl ce; do
      BA_RING_TRANSPORT=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N \
        --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --seq 65536 --steps 5 --warmup 3 --no-e2e \
        > gpurun_out/bench_n${N}_s65536_$tr.json 2> gpurun_out/bench_n${N}_s65536_$tr.err
      tail -1 gpurun_out/bench_n${N}_s65536_$tr.json
    done
  fi
fi
ls -la gpurun_out
