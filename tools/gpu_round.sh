#!/bin/bash
# One gpurun call = tests + bench (ours / torch baseline) + launch list + ncu full capture.
# Usage (on the GPU box, from the repo root):  bash tools/gpu_round.sh [tests] [bench] [launches] [ncu]
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
STAGES="${*:-tests bench launches ncu}"
GPUS="${GPUS:-1}"      # bench stages run on this many GPUs, launched exactly like the driver does (torchrun for N>1)
run_bench() {          # run_bench <extra args...>
  if [ "$GPUS" -gt 1 ]; then
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus "$GPUS" "$@"
  else
    timeout 300 python bench.py --gpus 1 "$@"
  fi
}
has() { [[ " $STAGES " == *" $1 "* ]]; }
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi_start.csv 2>&1

if has diag; then
  # numerics of the newest kernels first; a failing (or hanging: watchdog trap / timeout) tensor-core conv1 falls back to
  # the SIMT kernels for the rest of the call so the other measurements stay meaningful
  timeout 180 python tools/gpu_diag_lenet.py ${DIAG:-conv1 fc1_dgrad} > gpurun_out/diag.log 2>&1
  echo "diag exit=$?"; grep -E "FAIL|EXCEPTION|Error|error" gpurun_out/diag.log | head -20; grep -c " ok" gpurun_out/diag.log
  if grep -qE "conv1_(fwd|wgrad)_tc.*FAIL|check_conv1.*EXCEPTION" gpurun_out/diag.log; then
    echo "conv1 tensor-core kernels NOT healthy -> DMNIST_CONV1_TC=0 for the rest of this call"
    export DMNIST_CONV1_TC=0
  fi
fi

if has mgtests; then
  # multi-GPU boxes are charged per GPU: only the multi-rank tests there
  timeout 600 python -m pytest tests/test_fused_sync_gpu.py -m gpu -x -q > gpurun_out/pytest_multigpu.log 2>&1
  echo "pytest(multi-gpu) exit=$?" >> gpurun_out/pytest_multigpu.log
  tail -4 gpurun_out/pytest_multigpu.log
fi

if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi

if has bench; then
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
      --format=csv -lms 200 > gpurun_out/clocks.csv 2>&1 &
  SMI=$!
  run_bench --steps 300 --warmup 20 --kernel-times > gpurun_out/bench_ours_$GPUS.json 2> gpurun_out/bench_ours_$GPUS.err
  echo "bench ours exit=$?"; cat gpurun_out/bench_ours_$GPUS.json; grep KERNEL_TIMES gpurun_out/bench_ours_$GPUS.err; tail -3 gpurun_out/bench_ours_$GPUS.err
  if has torch; then
    run_bench --impl torch_ddp --steps 300 --warmup 20 > gpurun_out/bench_torch_$GPUS.json 2> gpurun_out/bench_torch_$GPUS.err
    echo "bench torch exit=$?"; cat gpurun_out/bench_torch_$GPUS.json
    timeout 300 python bench.py --impl reference --gpus 1 > gpurun_out/bench_ref_1.json 2>&1
  fi
  kill $SMI
fi

if has kofn; then  # (mlp3 only when MLP3=1)
  # backup-worker configuration of BASELINE.json: K = N-2 of N with one replica delayed on the device every step
  K=$((GPUS-2)); [ "$K" -lt 1 ] && K=1
  run_bench --k $K --straggler $((GPUS-1)):1.0:300 --steps 200 --warmup 10 > gpurun_out/bench_kofn_$GPUS.json 2> gpurun_out/bench_kofn_$GPUS.err
  echo "bench K=$K of $GPUS (rank $((GPUS-1)) delayed 300 us/step) exit=$?"; cat gpurun_out/bench_kofn_$GPUS.json
  if [ "${MLP3:-0}" = "1" ]; then
    run_bench --model mlp3 --batch 8192 --hidden 4096 --steps 30 --warmup 5 > gpurun_out/bench_mlp3_$GPUS.json 2> gpurun_out/bench_mlp3_$GPUS.err
    echo "bench mlp3 B=8192 exit=$?"; cat gpurun_out/bench_mlp3_$GPUS.json
  fi
fi

if has ab; then
  # A/B switches: one short bench per entry of $AB (e.g. AB="DMNIST_CONV1_FWD=1 DMNIST_PRIO=0"), ms_per_step only
  for kv in ${AB:-}; do
    r=$(env ${kv//,/ } bash -c "$(declare -f run_bench); GPUS=$GPUS; run_bench --steps 200 --warmup 10" 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('sync_phases_ns'), d.get('sync_early_phases_ns'))")
    echo "AB $kv -> ms_per_step (device, e2e), sync phases = $r" | tee -a gpurun_out/ab.txt
  done
fi

if has trace; then
  run_bench --steps 20 --warmup 5 --trace gpurun_out/timeline > gpurun_out/trace_run.log 2>&1
  echo "trace exit=$?"; cat gpurun_out/timeline.txt
fi

if has sanitize; then
  # compute-sanitizer over the kernel numerics checks (the reference has no race/memory checking at all)
  for tool in memcheck racecheck; do
    timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/gpu_diag_lenet.py ${SAN:-conv1 fc2_loss fc1_dgrad conv2_fwd} \
        > gpurun_out/sanitizer_$tool.log 2>&1
    echo "compute-sanitizer $tool exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_$tool.log | head -8
  done
fi

if has probe; then
  timeout 120 python tools/gpu_probe_coresidency.py > gpurun_out/coresidency_probe.log 2>&1; cat gpurun_out/coresidency_probe.log | cut -c1-220
fi

if has dgradtl; then
  timeout 120 python tools/gpu_timeline_dgrad.py > gpurun_out/dgrad_timeline.txt 2>&1; cat gpurun_out/dgrad_timeline.txt
fi

if has sweep; then
  DM_SWEEP_MAX=${DM_SWEEP_MAX:-67108864} timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$GPUS" \
      --master-addr 127.0.0.1 --master-port 29519 tools/allreduce_sweep.py > gpurun_out/sweep_$GPUS.log 2>&1
  echo "sweep exit=$?"; tail -2 gpurun_out/sweep_$GPUS.log | cut -c1-2500
fi

if has nvlsab; then
  # same bench with NVLS off (P2P reduce / push): what the in-switch reduction buys
  DMNIST_NVLS=0 run_bench --steps 300 --warmup 20 > gpurun_out/bench_ours_${GPUS}_p2p.json 2> gpurun_out/bench_ours_${GPUS}_p2p.err
  echo "bench ours P2P exit=$?"; cat gpurun_out/bench_ours_${GPUS}_p2p.json
fi

if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/launches_run.log 2>&1
  echo "ncu launches exit=$?"
fi

if has ncu; then
  # two whole training steps of our kernels, full metric set, source-correlated
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:'conv1_|conv2_|gemm_tc|fc2_loss|unpool2|fused_sync' -s 36 -c 24 \
      -f -o gpurun_out/lenet_step python bench.py --steps 3 --warmup 3 --no-graph > gpurun_out/ncu_full_run.log 2>&1
  echo "ncu full exit=$?"
  ls -la gpurun_out/
fi
