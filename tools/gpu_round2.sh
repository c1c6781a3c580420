#!/bin/bash
# Round-2 GPU stages (one gpurun call = several of them).  Usage on the box:  [GPUS=n] bash tools/gpu_round2.sh <stage>...
#   diag      numerics self-checks named in $DIAG (python tools/gpu_diag_lenet.py)
#   c1wg64    conv1_wgrad 64-pixel-tile variant: numerics, then A/B bench
#   tests     pytest -m gpu (all);   mgtests  only the multi-GPU aggregation tests
#   bench     bench.py ours (300 steps) + --kernel-times;   base  torch+NCCL baseline eager + CUDA-graphed
#   ab        one short bench per entry of $AB (comma = several variables in one entry)
#   trace     CUPTI timeline of graph-replayed steps
#   entry     throughput of the reference-compatible entrypoint (src/mnist_distributed_train.py) from its own log lines
#   kofn      K = N-2 of N with a device-side straggler;   mlp3  3-layer MLP batch 8192;   sweep  allreduce sweep
#   cdf       cdf-mode run with an injected straggler -> ELAPSED TIMES / time_cdfs.png
set -u
mkdir -p gpurun_out
STAGES="$*"
GPUS="${GPUS:-1}"
has() { [[ " $STAGES " == *" $1 "* ]]; }
run_py() {   # run_py <timeout> <script> args...   (torchrun for N > 1, exactly like the driver)
  local t=$1; shift
  if [ "$GPUS" -gt 1 ]; then
    timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 --master-port 29517 "$@"
  else
    timeout "$t" python "$@"
  fi
}
run_bench() { run_py 400 bench.py --gpus "$GPUS" "$@"; }
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi_start.csv 2>&1

if has diag; then
  timeout 300 python tools/gpu_diag_lenet.py ${DIAG:-bucketed end_to_end} > gpurun_out/diag.log 2>&1
  echo "diag exit=$?"; grep -E "FAIL|EXCEPTION|Error" gpurun_out/diag.log | head -20; sed -n '/^====/,$p' gpurun_out/diag.log | cut -c1-150
fi
if has c1wg64; then
  DMNIST_C1WG_TILE=64 timeout 200 python tools/gpu_diag_lenet.py conv1_wgrad end_to_end > gpurun_out/diag_c1wg64.log 2>&1
  echo "c1wg64 diag exit=$?"; sed -n '/^====/,$p' gpurun_out/diag_c1wg64.log | cut -c1-150
fi
if has mgtests; then
  timeout 900 python -m pytest tests/test_fused_sync_gpu.py -m gpu -x -q ${MGTESTS_K:+-k "$MGTESTS_K"} > gpurun_out/pytest_multigpu.log 2>&1
  echo "pytest(multi-gpu) exit=$?" >> gpurun_out/pytest_multigpu.log; tail -30 gpurun_out/pytest_multigpu.log | cut -c1-400
fi
if has tests; then
  timeout ${TESTS_TIMEOUT:-1200} python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log | cut -c1-400
fi
if has bench; then
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
      --format=csv -lms 200 > gpurun_out/clocks_$GPUS.csv 2>&1 &
  SMI=$!
  run_bench --steps 300 --warmup 20 ${BENCH_ARGS:-} > gpurun_out/bench_ours_$GPUS.json 2> gpurun_out/bench_ours_$GPUS.err
  echo "bench ours exit=$?"; cat gpurun_out/bench_ours_$GPUS.json; grep -m1 KERNEL_TIMES gpurun_out/bench_ours_$GPUS.err; tail -3 gpurun_out/bench_ours_$GPUS.err
  if [ "${BENCH_K20:-1}" = "1" ]; then
    # the driver's own invocation: 20 timed steps
    run_bench --steps 20 --warmup 5 > gpurun_out/bench_ours_${GPUS}_k20.json 2>> gpurun_out/bench_ours_$GPUS.err
    echo "bench ours (20 steps, as the driver runs it) exit=$?"; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_ours_${GPUS}_k20.json').read().strip().splitlines()[-1]); print('k20 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
  fi
  kill $SMI
fi
if has base; then
  run_bench --impl torch_ddp --steps 300 --warmup 20 > gpurun_out/bench_torch_$GPUS.json 2> gpurun_out/bench_torch_$GPUS.err
  echo "bench torch eager exit=$?"; cat gpurun_out/bench_torch_$GPUS.json
  run_bench --impl torch_ddp --graphed --steps 300 --warmup 20 > gpurun_out/bench_torchgraph_$GPUS.json 2> gpurun_out/bench_torchgraph_$GPUS.err
  echo "bench torch CUDA-graphed exit=$?"; cat gpurun_out/bench_torchgraph_$GPUS.json; tail -3 gpurun_out/bench_torchgraph_$GPUS.err
fi
if has ab; then
  for kv in ${AB:-}; do
    r=$(env ${kv//,/ } bash -c "$(declare -f run_py run_bench); GPUS=$GPUS; run_bench --steps 300 --warmup 20" 2>gpurun_out/ab_last.err | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['e2e']['ms_per_step']*1e3,2), d.get('sync_phases_ns'), d.get('sync_early_phases_ns'))")
    echo "AB $kv -> us/step (device, e2e), sync phases = $r" | tee -a gpurun_out/ab_$GPUS.txt
  done
fi
if has trace; then
  env ${TRACE_ENV:-} bash -c "$(declare -f run_py run_bench); GPUS=$GPUS; run_bench --steps 20 --warmup 5 --trace gpurun_out/timeline_$GPUS" > gpurun_out/trace_run.log 2>&1
  echo "trace exit=$?"; cat gpurun_out/timeline_$GPUS.txt
fi
if has entry; then
  for le in 1 100; do
    rm -rf /tmp/entry_train
    run_py 300 src/mnist_distributed_train.py --job_name=worker --batch_size=256 --max_steps=3000 --log_every=$le \
        --train_dir=/tmp/entry_train --save_interval_secs=1000 --initial_learning_rate=0.01 > gpurun_out/entry_${GPUS}_le$le.log 2>&1
    echo "entry log_every=$le exit=$?"
    python - <<PY
import re
t=open("gpurun_out/entry_${GPUS}_le$le.log").read()
ex=[float(x) for x in re.findall(r"\(([0-9.]+) examples/sec;", t)]
el=re.findall(r"Elapsed Time: ([0-9.]+)", t)
ex=ex[len(ex)//5:]
if ex:
    ex.sort(); print("entry N=$GPUS log_every=$le: median examples/sec/replica = %.0f (%.1f us/step), lines=%d, elapsed=%s" % (ex[len(ex)//2], 256e6/ex[len(ex)//2], len(ex), el))
else:
    print("no examples/sec lines", t[-1500:])
PY
  done
fi
if has kofn; then
  K=$((GPUS-2)); [ "$K" -lt 1 ] && K=1
  DMNIST_DEBUG_SYNC=1 DMNIST_BENCH_ABORT_S=60 run_bench --k $K --straggler $((GPUS-1)):1.0:300 --steps 200 --warmup 10 > gpurun_out/bench_kofn_$GPUS.json 2> gpurun_out/bench_kofn_$GPUS.err
  echo "bench K=$K of $GPUS (rank $((GPUS-1)) delayed 300 us/step) exit=$?"; cat gpurun_out/bench_kofn_$GPUS.json; tail -5 gpurun_out/bench_kofn_$GPUS.err
fi
if has mlp3; then
  run_bench --model mlp3 --batch 8192 --hidden 4096 --steps 30 --warmup 5 > gpurun_out/bench_mlp3_$GPUS.json 2> gpurun_out/bench_mlp3_$GPUS.err
  echo "bench mlp3 B=8192 exit=$?"; cat gpurun_out/bench_mlp3_$GPUS.json; tail -3 gpurun_out/bench_mlp3_$GPUS.err
fi
if has sweep; then
  for nv in ${SWEEP_NVLS:-1 0}; do
    DMNIST_NVLS=$nv DM_SWEEP_MAX=${DM_SWEEP_MAX:-268435456} run_py 400 tools/allreduce_sweep.py > gpurun_out/sweep_${GPUS}_nvls$nv.log 2>&1
    echo "sweep N=$GPUS NVLS=$nv exit=$?"; tail -3 gpurun_out/sweep_${GPUS}_nvls$nv.log | cut -c1-3000
  done
fi
if has cdf; then
  # cdf mode (full barrier + per-iteration timing) with a device-side straggler on the last rank: the ELAPSED TIMES tables come
  # from the %globaltimer stamps of the step's kernels, so only the delayed rank shows a tail
  rm -rf /tmp/cdf_train; mkdir -p gpurun_out/cdf_$GPUS
  run_py ${CDF_TIMEOUT:-300} src/mnist_distributed_train.py --job_name=worker --batch_size=256 --max_steps=520 --log_every=100 \
      --worker_times_cdf_method=true --inject_straggler=$((GPUS-1)):0.3:300 --train_dir=/tmp/cdf_train --save_interval_secs=1000 \
      > gpurun_out/cdf_$GPUS/b256_straggler_${GPUS}gpu_out_master 2>&1
  echo "cdf exit=$?"; grep -c "ELAPSED TIMES" gpurun_out/cdf_$GPUS/b256_straggler_${GPUS}gpu_out_master
  python - <<PY
import sys
sys.path.insert(0, "tools")
import benchmark as B
d = "gpurun_out/cdf_$GPUS"
print(B.plot_time_cdfs(d, d))
import collections
ct = B.extract_compute_times(d + "/b256_straggler_${GPUS}gpu_out_master")
per = collections.defaultdict(list)
for t, w, it in ct:
    per[w].append(t)
for w in sorted(per):
    v = sorted(per[w]); n = len(v)
    print("worker %d: n=%d median=%.1f us p95=%.1f us max=%.1f us" % (w, n, v[n // 2] * 1e6, v[int(n * 0.95)] * 1e6, v[-1] * 1e6))
PY
fi
if has ncu; then
  # one full-set capture of every kernel of the (eager, un-graphed) step on ONE GPU; read here with tools/ncu_summary.py
  timeout ${NCU_TIMEOUT:-900} ncu --set full --clock-control none --import-source on \
      -k regex:'conv1_|conv2_|gemm_tc|fc2_|bucket_' -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-36} \
      -f -o gpurun_out/lenet_step_r2 python bench.py --steps 3 --warmup 3 --no-graph > gpurun_out/ncu_full_run.log 2>&1
  echo "ncu full exit=$?"; ls -la gpurun_out/*.ncu-rep
fi
if has sanitize; then
  for tool in memcheck racecheck; do
    timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/gpu_diag_lenet.py ${SAN:-conv1 fc2_loss fc1_dgrad conv2_fwd bucketed_step} \
        > gpurun_out/sanitizer_$tool.log 2>&1
    echo "compute-sanitizer $tool exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_$tool.log | head -8
  done
fi
if has matrix; then
  # the reference's experiment matrices on this box: K-of-N sweep + interval sweep (time-to-accuracy, step rate, loss) and the
  # compute-time CDF study; figures + scraped logs land in gpurun_out/matrix_*
  if [ -n "${MATRIX_FILES:-}" ]; then
    # explicit list of configuration files (budgeted runs)
    timeout ${MATRIX_TIMEOUT:-600} python tools/benchmark.py select_files $MATRIX_FILES --n_iters=${MATRIX_ITERS:-19000} \
        --outdir=gpurun_out/matrix_sel --dest=gpurun_out/matrix_sel > gpurun_out/matrix_sel.log 2>&1
    echo "matrix (selected files) exit=$?"; grep -E "Currently on iteration|timeout waiting|Traceback|Error" gpurun_out/matrix_sel.log | head -40
    grep -A14 "^File:" gpurun_out/matrix_sel.log | head -40; ls gpurun_out/matrix_sel | head -60
    MATRIX=""
  fi
  for m in ${MATRIX-8_gpus time_cdf_cfgs}; do
    files=""
    for f in $(ls cfg/$m); do
      # MATRIX_SKIP: space-separated substrings of configuration names to leave out (budget)
      skip=0; for pat in ${MATRIX_SKIP:-}; do [[ "$f" == *"$pat"* ]] && skip=1; done
      [ "$skip" = 0 ] && files="$files cfg/$m/$f"
    done
    timeout ${MATRIX_TIMEOUT:-600} python tools/benchmark.py select_files $files --n_iters=${MATRIX_ITERS:-19000} \
        --outdir=gpurun_out/matrix_$m --dest=gpurun_out/matrix_$m > gpurun_out/matrix_$m.log 2>&1
    echo "matrix $m exit=$?"; grep -E "Currently on iteration|timeout waiting|Traceback|Error" gpurun_out/matrix_$m.log | head -40; ls gpurun_out/matrix_$m | head -60
  done
fi
ls gpurun_out | head -50
