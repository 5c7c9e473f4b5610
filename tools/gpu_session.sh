#!/bin/bash
# One GPU-box session that measures as much as possible (GPU slots are scarce): every stage logs to
# gpurun_out/<tag>_*.{log,json,csv} and never aborts the following stages.
#   tools/gpu_session.sh <tag> [stages...]      stages: tests bench bf16 workloads ncu sanitize
tag=${1:-s}; shift
stages=${@:-tests bench}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# the snapshot may have been taken between a source edit and its rebuild: build() is digest-stamped (no-op when current)
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/${tag}_build.log 2>&1 || tail -5 gpurun_out/${tag}_build.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
for st in $stages; do
  case $st in
    tests)
      timeout 900 python -m pytest tests/test_reference_boundary.py -q -m gpu > gpurun_out/${tag}_boundary.log 2>&1
      tail -5 gpurun_out/${tag}_boundary.log
      timeout 1500 python -m pytest tests -q -m gpu --durations=12 -p no:cacheprovider > gpurun_out/${tag}_gputests.log 2>&1
      tail -25 gpurun_out/${tag}_gputests.log ;;
    quicktests)
      timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gemm_tc or mlp_chain or trajectory_tf32x3 or criteo_shape" -p no:cacheprovider > gpurun_out/${tag}_quick.log 2>&1
      tail -25 gpurun_out/${tag}_quick.log ;;
    bench)
      timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      tail -c 1500 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err ;;
    bf16)
      for prec in bf16 tf32 fp32; do
        timeout 300 python bench.py --precision $prec --steps-only > gpurun_out/${tag}_bench_$prec.json 2> gpurun_out/${tag}_bench_$prec.err
        cat gpurun_out/${tag}_bench_$prec.json; tail -2 gpurun_out/${tag}_bench_$prec.err
      done ;;
    workloads)
      for w in dcnv2 din xdeepfm; do
        timeout 400 python bench.py --workload $w --steps 100 --cpu-seconds 6 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
        head -c 600 gpurun_out/${tag}_bench_$w.json; echo; tail -2 gpurun_out/${tag}_bench_$w.err
      done ;;
    debug)
      for m in tf32 tf32x3 bf16; do timeout 120 python tools/debug_chain.py $m 2>&1 | tail -8; done
      # PDL A/B on the default step, the 128x160 3xTF32 tile on C3, and the full GPU suite
      timeout 200 python bench.py --steps-only > gpurun_out/${tag}_pdl0.json 2> gpurun_out/${tag}_pdl0.err; cat gpurun_out/${tag}_pdl0.json
      B2_PDL=1 timeout 200 python bench.py --steps-only > gpurun_out/${tag}_pdl1.json 2> gpurun_out/${tag}_pdl1.err; cat gpurun_out/${tag}_pdl1.json; tail -3 gpurun_out/${tag}_pdl1.err
      timeout 300 python bench.py --workload dcnv2 --steps 100 --no-cpu-baseline > gpurun_out/${tag}_bench_dcnv2.json 2> gpurun_out/${tag}_bench_dcnv2.err; head -c 400 gpurun_out/${tag}_bench_dcnv2.json; echo
      timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_gputests.log 2>&1; tail -8 gpurun_out/${tag}_gputests.log ;;
    ab)
      # A/B of the launch options on the default step: TMA-store epilogue on/off, inline 3xTF32 split on/off
      for v in "B2_X3_CHAIN_KB=0" "B2_X3_CHAIN_KB=64" "B2_X3_CHAIN_KB=16" "B2_X3_CHAIN_KB=64 B2_X3_BN_MAX=256"; do
        n=$(echo $v | tr -d ' =_A-Z')
        env $v timeout 200 python bench.py --steps-only > gpurun_out/${tag}_ab_$n.json 2> gpurun_out/${tag}_ab_$n.err
        echo "$v: $(python -c "import json,sys; d=json.loads(open('gpurun_out/${tag}_ab_$n.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
      done ;;
    kernels)
      timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${tag}_kern.json 2> gpurun_out/${tag}_kern.err
      python -c "
import json; d=json.loads(open('gpurun_out/${tag}_kern.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['clocks'])
for k,v in d['kernels'].items(): print(k, round(v['ms']*1e3,2), 'us', v.get('frac_of_tf32_peak_counting_passes', v.get('frac_of_measured_hbm')))
print({k:(v['launches'], round(v['us'],1)) for k,v in d['step_profile']['calls'].items()})
" 2>&1 | tail -14; tail -3 gpurun_out/${tag}_kern.err ;;
    probe)
      timeout 300 python tools/gemm_probe.py > gpurun_out/${tag}_probe.log 2>&1; cat gpurun_out/${tag}_probe.log | cut -c1-400 ;;
    trace)
      for k in fwd dgrad wgrad; do for m in aux inline; do timeout 120 python tools/gemm_trace.py $m $k > gpurun_out/${tag}_trace_${m}_$k.log 2>&1; done; done; cat gpurun_out/${tag}_trace_aux_fwd.log gpurun_out/${tag}_trace_aux_dgrad.log ;;
    pdltests)
      B2_PDL=1 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x --deselect tests/test_reference_boundary.py > gpurun_out/${tag}_pdltests.log 2>&1; tail -5 gpurun_out/${tag}_pdltests.log ;;
    dlrm_small)
      timeout 300 python bench.py --workload dlrm --vocab-scale 0.01 --steps 30 --warmup 5 --nbatches 8 --steps-only > gpurun_out/${tag}_bench_dlrmsmall_n1.json 2> gpurun_out/${tag}_bench_dlrmsmall_n1.err
      cat gpurun_out/${tag}_bench_dlrmsmall_n1.json; tail -30 gpurun_out/${tag}_bench_dlrmsmall_n1.err ;;
    dlrm1)
      timeout 600 python bench.py --workload dlrm --steps 20 --warmup 5 --nbatches 8 > gpurun_out/${tag}_bench_dlrm_n1.json 2> gpurun_out/${tag}_bench_dlrm_n1.err
      head -c 800 gpurun_out/${tag}_bench_dlrm_n1.json; echo; tail -3 gpurun_out/${tag}_bench_dlrm_n1.err ;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 300 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --steps 4 --warmup 3 --graph 0 --steps-only --nbatches 4 > gpurun_out/${tag}_ncu_launch.log 2>&1
      tail -3 gpurun_out/${tag}_ncu_launch.log
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 12 -c 9 -o gpurun_out/${tag}_gemm \
        python bench.py --steps 1 --warmup 3 --graph 0 --steps-only --nbatches 4 > gpurun_out/${tag}_ncu_gemm.log 2>&1
      tail -3 gpurun_out/${tag}_ncu_gemm.log ;;
    sanitize)
      timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q -m gpu -x -p no:cacheprovider \
        -k "gemm_tc_mn_major and (128-32-32 or 76-44-36 or 132-260) or fused_backward or mlp_chain and 40 or virtual_ranks or shard_roundtrip or cin_fused or din_softmax or hot_row" \
        > gpurun_out/${tag}_memcheck.log 2>&1
      tail -8 gpurun_out/${tag}_memcheck.log ;;
    racecheck)
      # shared-memory hazards: the GEMM's patches / converter tiles, the push kernel's block list, CIN / DIN tiles
      timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -q -m gpu -x -p no:cacheprovider \
        -k "gemm_tc_mn_major and (128-32-32 or 76-44-36) or mlp_chain and 40 or virtual_ranks and 2 or cin_fused or din_softmax or hot_row" \
        > gpurun_out/${tag}_racecheck.log 2>&1
      tail -8 gpurun_out/${tag}_racecheck.log ;;
    c5prop)
      timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k c5_full_shape -p no:cacheprovider > gpurun_out/${tag}_c5prop.log 2>&1; tail -5 gpurun_out/${tag}_c5prop.log ;;
    ncu_cin)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:cin_ -s 18 -c 9 -o gpurun_out/${tag}_cin \
        python bench.py --workload xdeepfm --steps 1 --warmup 3 --graph 0 --steps-only --nbatches 4 > gpurun_out/${tag}_ncu_cin.log 2>&1
      tail -3 gpurun_out/${tag}_ncu_cin.log ;;
  esac
done
echo "[gpu_session $tag] done"
