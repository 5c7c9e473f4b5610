#!/bin/bash
# Multi-GPU session (charged N x the box time: keep it short).  usage: tools/gpu_session_multi.sh <tag> <N> [stages...]
#   stages: parity profile deepfm dlrm dlrm_small, each optionally as stage@ranks (e.g. deepfm@4)
tag=$1; N=$2; shift; shift
stages=${@:-parity profile deepfm}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# the snapshot may have been taken between a source edit and its rebuild: build() is digest-stamped (no-op when current)
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/${tag}_build.log 2>&1 || tail -5 gpurun_out/${tag}_build.log
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "${@:2}"; }
N0=$N
for tok in $stages; do
  st=${tok%@*}; N=$N0
  if [[ $tok == *@* ]]; then N=${tok#*@}; fi      # stage@ranks: run this stage on fewer ranks of the same box
  case $st in
    parity)
      for model in ${PARITY_MODELS:-DeepFM DLRM}; do
        run 300 tools/dist_sharded_check.py --model $model --precision tf32x3 > gpurun_out/${tag}_parity_${model}_n$N.log 2>&1
        grep -E "err|OK|FAIL|Error" gpurun_out/${tag}_parity_${model}_n$N.log | tail -$((N + 3))
      done ;;
    profile)
      run 300 tools/profile_sharded.py > gpurun_out/${tag}_phases_n$N.log 2>&1
      tail -2 gpurun_out/${tag}_phases_n$N.log | cut -c1-1500 ;;
    deepfm)
      run 400 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/${tag}_bench_deepfm_n$N.json 2> gpurun_out/${tag}_bench_deepfm_n$N.err
      head -c 700 gpurun_out/${tag}_bench_deepfm_n$N.json; echo; tail -2 gpurun_out/${tag}_bench_deepfm_n$N.err ;;
    dlrm)
      run 600 bench.py --gpus $N --workload dlrm --steps 30 --warmup 5 --nbatches 8 > gpurun_out/${tag}_bench_dlrm_n$N.json 2> gpurun_out/${tag}_bench_dlrm_n$N.err
      head -c 900 gpurun_out/${tag}_bench_dlrm_n$N.json; echo; tail -3 gpurun_out/${tag}_bench_dlrm_n$N.err ;;
    dlrm_small)
      run 300 bench.py --gpus $N --workload dlrm --vocab-scale 0.01 --steps 30 --warmup 5 --nbatches 8 --steps-only > gpurun_out/${tag}_bench_dlrmsmall_n$N.json 2> gpurun_out/${tag}_bench_dlrmsmall_n$N.err
      cat gpurun_out/${tag}_bench_dlrmsmall_n$N.json; tail -3 gpurun_out/${tag}_bench_dlrmsmall_n$N.err ;;
  esac
done
echo "[gpu_session_multi $tag N=$N] done"
