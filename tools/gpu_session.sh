#!/bin/bash
# One gpurun call that runs everything still waiting for a B200 (see DESIGN.md "Open items").  Writes to gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh lu ncu'            (1 GPU)
#   gpurun --gpus 4 --timeout 1200 -- 'bash tools/gpu_session.sh dist4'    (4 GPUs)
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_session.sh dist8'    (8 GPUs)
# Sections: f32 (FP32 tcgen05 GEMM: parity, speed, ncu), sanitize (compute-sanitizer memcheck / racecheck on the local kernels), e2e1 / e2eN (end-to-end knobs added without a GPU: early C download, upload policy, panel count), lu (LU seam tests + LU bench host vs GPU), pending1 (redistribution / Yamamoto tests + their measurements), ncu (full capture at n=32768 for roofline.traffic, pack kernels),
#           dist4 (second-pass parity cases, update_A with T), dist8 (fused depth sum on 2x2x2, skip_unused_uploads)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REF=oracle/_ref
for section in "$@"; do
  case "$section" in
    lu)
      # xfail(strict=False) cases report XPASS when they pass; -rA lists every outcome
      timeout 1200 python -m pytest tests/test_zz_lu_offload_gpu.py -m gpu -q -rA -p no:cacheprovider \
        > gpurun_out/lu_offload_pytest.log 2>&1
      tail -40 gpurun_out/lu_offload_pytest.log
      if [ -x $REF/lu_bench_host ] && [ -x $REF/dropin/lu_bench_gpu ]; then
        for n in 4096 8192; do
          echo "== LU bench n=$n host fallback" >> gpurun_out/lu_bench.log
          timeout 600 $REF/mpirun -np 4 -timeout 500 $REF/lu_bench_host -n $n -b_sm 64 -b_lrg 512 -num_iter 2 >> gpurun_out/lu_bench.log 2>&1
          echo "== LU bench n=$n GPU seam" >> gpurun_out/lu_bench.log
          timeout 600 $REF/mpirun -np 4 -timeout 500 $REF/dropin/lu_bench_gpu -n $n -b_sm 64 -b_lrg 512 -num_iter 2 >> gpurun_out/lu_bench.log 2>&1
        done
        grep -E "==|Gigaflops|elapsed|failed" gpurun_out/lu_bench.log
      fi
      ;;
    ncu)
      # one full capture of the headline launch (n = 32768, ~2 s per pass) for roofline.traffic, and of the pack kernels
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -c 1 \
        -o gpurun_out/ncu_full_gemm_n32768 -f tools/gemm_probe speed 32768 1 > gpurun_out/ncu_full_gemm_n32768.log 2>&1
      python tools/ncu_summary.py gpurun_out/ncu_full_gemm_n32768.ncu-rep gpurun_out/ncu_full_gemm_n32768.csv | tail -3
      timeout 900 ncu --set full --clock-control none -k regex:'lda_|transpose|sparse_rows' -c 12 \
        -o gpurun_out/ncu_full_pack -f tools/gemm_probe pack 16384 > gpurun_out/ncu_full_pack.log 2>&1
      python tools/ncu_summary.py gpurun_out/ncu_full_pack.ncu-rep gpurun_out/ncu_full_pack.csv | tail -14
      ;;
    f32)
      # the FP32 tcgen05 GEMM: parity (first contact with hardware), first speeds, and an ncu pass that shows the tensor pipe
      timeout 600 python tests/f32_worker.py --bench > gpurun_out/f32_worker.json 2> gpurun_out/f32_worker.err
      tail -3 gpurun_out/f32_worker.json; tail -5 gpurun_out/f32_worker.err
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f32_umma -c 1 -f -o gpurun_out/f32_gemm \
        python tests/f32_worker.py --one 8192 > gpurun_out/f32_ncu.log 2>&1
      ncu -i gpurun_out/f32_gemm.ncu-rep --page raw --csv > gpurun_out/f32_gemm_raw.csv 2>/dev/null
      ;;
    pending1)
      # redistribution / Yamamoto tests on one GPU and the not-yet-measured widening rows
      timeout 1200 python -m pytest tests/test_zz_redist_gpu.py -m gpu -q -rA -p no:cacheprovider > gpurun_out/redist_pytest.log 2>&1
      tail -25 gpurun_out/redist_pytest.log
      # the hot kernel with B chunk-major through one tensor map: parity (bit for bit against the plain launch), then one merged
      # launch against seven per-chunk launches at b = 8192 / 16384 (the saving candmc_set_merge_last_panel is after)
      timeout 600 python tests/bchunk_worker.py --bench > gpurun_out/bchunk_worker.json 2> gpurun_out/bchunk_worker.err
      tail -2 gpurun_out/bchunk_worker.json; tail -3 gpurun_out/bchunk_worker.err
      timeout 900 python tools/bench_configs.py --pending > gpurun_out/bench_pending_1gpu.jsonl 2> gpurun_out/bench_pending_1gpu.err
      cat gpurun_out/bench_pending_1gpu.jsonl
      ;;
    dist4)
      CANDMC_TEST_VERBOSE=1 timeout 1100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_worker.py > gpurun_out/dist4_worker.log 2>&1
      tail -15 gpurun_out/dist4_worker.log
      CANDMC_TEST_PENDING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29535 tests/dist_worker.py > gpurun_out/dist4_pending.log 2>&1
      tail -8 gpurun_out/dist4_pending.log
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29536 \
        tools/bench_configs.py --pending > gpurun_out/bench_pending_4gpu.jsonl 2> gpurun_out/bench_pending_4gpu.err
      cat gpurun_out/bench_pending_4gpu.jsonl
      # config 2 (b = 8192) with 2048-deep k-chunks instead of 1024: per-chunk epilogue and tail-wave losses halve
      CANDMC_MIN_KCHUNK=2048 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29541 tools/bench_configs.py > gpurun_out/bench_configs_4gpu_kc2048.jsonl 2> gpurun_out/bench_configs_4gpu_kc2048.err
      cat gpurun_out/bench_configs_4gpu_kc2048.jsonl
      # the last panel of a sweep in one launch over its k-chunks (opt-in): parity of the validated suite with it, then the headline
      # grid and configs 2 / 4 / 5 (config 2 is where per-chunk epilogues and tails cost most)
      CANDMC_TEST_MERGE_PANELS=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29542 tests/dist_worker.py > gpurun_out/dist4_merge.log 2>&1
      tail -3 gpurun_out/dist4_merge.log
      for knobs in "" "--merge-panels 1" "--merge-panels 2" "--merge-panels 3"; do
        echo "== bench 4 GPUs --no-e2e $knobs" >> gpurun_out/dist4_bench_merge.log
        timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 \
          bench.py --gpus 4 --steps 5 --warmup 3 --no-e2e $knobs >> gpurun_out/dist4_bench_merge.log 2>&1
      done
      grep -E "==|\"metric\"" gpurun_out/dist4_bench_merge.log | cut -c1-400
      CANDMC_MERGE_PANELS=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29544 tools/bench_configs.py > gpurun_out/bench_configs_4gpu_merge.jsonl 2> gpurun_out/bench_configs_4gpu_merge.err
      cat gpurun_out/bench_configs_4gpu_merge.jsonl
      # the opt-in peer-memory paths: parity first (same worker, switches from the environment), then configs 2 / 4 / 5 with
      # SUMMA panels and Cannon shifts on copy engines
      CANDMC_TEST_PANEL_TRANSPORT=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29538 tests/dist_worker.py > gpurun_out/dist4_transport.log 2>&1
      tail -4 gpurun_out/dist4_transport.log
      CANDMC_PANEL_TRANSPORT=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29539 tools/bench_configs.py > gpurun_out/bench_configs_4gpu_transport.jsonl 2> gpurun_out/bench_configs_4gpu_transport.err
      cat gpurun_out/bench_configs_4gpu_transport.jsonl
      ;;
    e2e1)
      # one GPU: the host-streamed multiply with 8 equal panels (measured), the graduated cut (default, never measured: the
      # plain run reports both passes in e2e_passes), 16 and 32 equal panels
      for knobs in "" "--host-panels 16" "--host-panels 32"; do
        echo "== bench 1 GPU $knobs" >> gpurun_out/e2e1_bench.log
        timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline $knobs >> gpurun_out/e2e1_bench.log 2>&1
      done
      grep -E "==|\"metric\"" gpurun_out/e2e1_bench.log | cut -c1-300
      ;;
    e2eN)
      # run under `gpurun --gpus N`: end-to-end leg with C leaving early (new default) vs one download at the end, and the old
      # upload-everything policy; N taken from the visible devices
      N=$(python -c "import torch; print(torch.cuda.device_count())")
      for knobs in "" "--late-c-download" "--upload-all-blocks --late-c-download" "--b-first-chunk-early"; do
        echo "== bench $N GPUs $knobs" >> gpurun_out/e2e${N}_bench.log
        timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29537 \
          bench.py --gpus $N --steps 3 --warmup 3 $knobs >> gpurun_out/e2e${N}_bench.log 2>&1
      done
      grep -E "==|\"metric\"" gpurun_out/e2e${N}_bench.log | cut -c1-300
      ;;
    dist8)
      CANDMC_TEST_PANEL_TRANSPORT=1 CANDMC_TEST_FUSED_GRIDS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
        --master-addr 127.0.0.1 --master-port 29540 tests/dist_worker.py > gpurun_out/dist8_peer_paths.log 2>&1
      tail -4 gpurun_out/dist8_peer_paths.log
      CANDMC_TEST_MERGE_PANELS=2 CANDMC_TEST_FUSED_GRIDS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
        --master-addr 127.0.0.1 --master-port 29546 tests/dist_worker.py > gpurun_out/dist8_merge.log 2>&1
      tail -3 gpurun_out/dist8_merge.log
      for knobs in "" "--merge-panels 2" "--merge-panels 3" "--fused-reduce 2" "--merge-panels 2 --fused-reduce 2" "--panel-transport" "--panel-transport --fused-reduce 2"; do
        echo "== bench 8 GPUs $knobs" >> gpurun_out/dist8_bench.log
        timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 \
          bench.py --gpus 8 --steps 5 --warmup 3 $knobs >> gpurun_out/dist8_bench.log 2>&1
      done
      # fewer, deeper k-chunks per panel: per-launch epilogue and tail losses against a longer exposed first broadcast
      for kc in 4096 8192; do
        echo "== bench 8 GPUs --min-kchunk $kc" >> gpurun_out/dist8_bench.log
        timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
          --master-port 29534 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --min-kchunk $kc >> gpurun_out/dist8_bench.log 2>&1
      done
      grep -E "==|\"metric\"" gpurun_out/dist8_bench.log | cut -c1-400
      ;;
    sanitize)
      # memcheck / racecheck over the local kernels at small sizes (FP64 GEMM all transposes, pack, FP32 GEMM)
      for tool in memcheck racecheck; do
        timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 tools/gemm_probe check \
          > gpurun_out/sanitize_${tool}_gemm_probe.log 2>&1; echo "$tool gemm_probe: exit $?"
        timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python tests/f32_worker.py --one 384 \
          > gpurun_out/sanitize_${tool}_f32.log 2>&1; echo "$tool f32: exit $?"
      done
      grep -h "ERROR SUMMARY" gpurun_out/sanitize_*.log
      ;;
    all1)
      # everything that needs one GPU, cheapest first contact first (about 25 minutes of box time)
      bash "$0" f32 pending1 lu e2e1 ncu
      ;;
    *) echo "unknown section $section";;
  esac
done
