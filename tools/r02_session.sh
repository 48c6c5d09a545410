#!/bin/bash
# Round-2 GPU sessions (one gpurun call each; everything lands in gpurun_out/r02_*).  Sections are independent and each
# step has its own timeout, so a hang costs minutes, not the call.
#   gpurun --timeout 1800 -- 'bash tools/r02_session.sh probe ncu1 f32 tests bench1 lu'     (session 1)
#   gpurun --timeout 900  -- 'bash tools/r02_session.sh pack'                                (session 2)
#   gpurun --timeout 900  -- 'bash tools/r02_session.sh tile'                                (session 4)
#   gpurun --timeout 900  -- 'bash tools/r02_session.sh seams'                               (NOT RUN: written after the budget was spent)
# The 4- and 8-GPU sessions are tools/r02_multi.py.  Keep gpurun_out/ under 64 MiB or nothing travels back (ncu_full
# deletes reports over 6 MB after summarising them).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NG=$(python -c "import torch; print(torch.cuda.device_count())" 2>/dev/null || echo 1)
trun() {   # trun <port> <script> args...   (one rank per visible GPU)
  local port=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $port "$@"
}
ncu_full() {   # ncu_full <out-stem> <kernel regex> <count> cmd...
  local stem=$1 rx=$2 cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -f -o $O/$stem "$@" > $O/$stem.log 2>&1
  python tools/ncu_summary.py $O/$stem.ncu-rep $O/$stem.csv > /dev/null 2>&1 && tail -n +3 $O/$stem.csv | cut -c1-400
  # gpurun_out/ only travels back when it stays under 64 MiB: keep the report itself only when it is small
  [ "$(stat -c %s $O/$stem.ncu-rep 2>/dev/null || echo 0)" -gt 6000000 ] && rm -f $O/$stem.ncu-rep
}
for section in "$@"; do
  echo "=== section $section ($(date +%T))"
  case "$section" in
    probe)
      # local kernels: the launches the sweeps are made of (beta = 1 chunks, reserved SMs, C-tile L2 prefetch on/off), square and
      # skinny shapes against the cuBLAS cross-check, pack kernels
      timeout 300 tools/gemm_probe chunk > $O/r02_gemm_probe_chunk.jsonl 2>&1; cat $O/r02_gemm_probe_chunk.jsonl | cut -c1-230
      timeout 300 tools/gemm_probe speed 1024 2048 4096 8192 16384 > $O/r02_gemm_probe_speed.jsonl 2>&1; cat $O/r02_gemm_probe_speed.jsonl | cut -c1-230
      timeout 120 tools/gemm_probe pack 8192 > $O/r02_pack_probe.jsonl 2>&1; timeout 120 tools/gemm_probe pack 16384 >> $O/r02_pack_probe.jsonl 2>&1
      cat $O/r02_pack_probe.jsonl
      ;;
    ncu1)
      # --set full captures at the shapes that matter: the beta = 1 chunk launch of the sweeps (146 SMs), a TN skinny launch,
      # the pack kernels, the headline 32768^3 launch (roofline.traffic)
      ncu_full r02_ncu_gemm_chunk_b1 gemm_f64_tma 1 tools/gemm_probe one N N 16384 16384 2048 1.0 2 0 1
      ncu_full r02_ncu_gemm_chunk_b1_pf gemm_f64_tma 1 tools/gemm_probe one N N 16384 16384 2048 1.0 2 1 1
      ncu_full r02_ncu_gemm_tn_skinny gemm_f64_tma 1 tools/gemm_probe one T N 512 8192 65536 0.0 0 0 1
      ncu_full r02_ncu_pack 'lda_|transpose' 9 tools/gemm_probe pack 8192
      ncu_full r02_ncu_gemm_n32768 gemm_f64_tma 1 tools/gemm_probe one N N 32768 32768 32768 0.0 0 0 1
      ;;
    f32)
      timeout 400 python tests/f32_worker.py --bench > $O/r02_f32_worker.json 2> $O/r02_f32_worker.err
      tail -4 $O/r02_f32_worker.json | cut -c1-600; tail -3 $O/r02_f32_worker.err
      ncu_full r02_ncu_gemm_f32 gemm_f32_umma 1 python tests/f32_worker.py --one 8192
      ;;
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q -rA -p no:cacheprovider > $O/r02_pytest_gpu_${NG}gpus.log 2>&1
      tail -60 $O/r02_pytest_gpu_${NG}gpus.log | grep -E "passed|failed|PASS|FAIL|XPASS|XFAIL|SKIP|Error" | cut -c1-200 | tail -70
      ;;
    bench1)
      timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/r02_bench_n32768_gpus1.json 2> $O/r02_bench_n32768_gpus1.err
      cut -c1-1500 $O/r02_bench_n32768_gpus1.json
      # launch list of the same command (per-launch times are cold-cache and serialised: shares, not absolutes)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches_bench_n32768.csv \
        python bench.py --gpus 1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_ncu_launches_bench.log 2>&1
      grep -c gemm_f64 $O/r02_ncu_launches_bench_n32768.csv
      ;;
    lu)
      REF=oracle/_ref
      if [ -x $REF/lu_bench_host ] && [ -x $REF/dropin/lu_bench_gpu ]; then
        for n in 4096 8192; do
          echo "== LU bench n=$n host fallback" >> $O/r02_lu_bench.log
          timeout 300 $REF/mpirun -np 4 -timeout 250 $REF/lu_bench_host -n $n -b_sm 64 -b_lrg 512 -num_iter 2 >> $O/r02_lu_bench.log 2>&1
          echo "== LU bench n=$n GPU seam" >> $O/r02_lu_bench.log
          timeout 300 $REF/mpirun -np 4 -timeout 250 $REF/dropin/lu_bench_gpu -n $n -b_sm 64 -b_lrg 512 -num_iter 2 >> $O/r02_lu_bench.log 2>&1
        done
        grep -E "==|Gigaflops|elapsed|failed|rror" $O/r02_lu_bench.log | head -20
      fi
      ;;
    seams)
      # the reference's own QR / LU / full -> band drivers over the integration/ seams (first B200 contact), then one timing each
      # of the reference's QR test all-host, with upd_A in the library and with cdgemm in the library (host matrices are
      # staged per call: this shows the drop-in, not config 5's speed)
      timeout 900 python -m pytest tests/test_zz_qr_dropin_gpu.py tests/test_zz_unseen_gpu.py tests/test_zz_aggregator_gpu.py -m gpu -q -rA \
        -p no:cacheprovider > $O/r02_seams_pytest.log 2>&1; tail -5 $O/r02_seams_pytest.log
      REF=oracle/_ref
      for exe in test_qr_2d_host dropin/test_qr_2d_gpu dropin/test_qr_2d_cdgemm_gpu; do
        if [ -x $REF/$exe ]; then
          echo "== $exe 4096 x 2048, b = 128, 1 rank" >> $O/r02_seams_qr.log
          ( time timeout 300 $REF/mpirun -np 1 -timeout 250 -threads 8 $REF/$exe 4096 2048 128 1 ) >> $O/r02_seams_qr.log 2>&1
        fi
      done
      grep -E "==|A-QR\|\|_2 =|real" $O/r02_seams_qr.log | head -12
      ;;
    tile)
      # A/B of the two CTA shapes of the DMMA GEMM (128 x 128 tiles, one CTA per SM / 128 x 64 tiles, two per SM)
      timeout 600 tools/gemm_probe tile > $O/r02_gemm_probe_tile.jsonl 2>&1; cut -c1-200 $O/r02_gemm_probe_tile.jsonl
      ;;
    pack)
      # the rewritten pack kernels: bit-exact tests first, then GB/s, then one ncu --set full pass over every pack kernel
      timeout 600 python -m pytest tests/test_local_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
      timeout 120 tools/gemm_probe pack 8192 > $O/r02_pack_probe.jsonl 2>&1; timeout 120 tools/gemm_probe pack 16384 >> $O/r02_pack_probe.jsonl 2>&1
      cat $O/r02_pack_probe.jsonl
      timeout 600 ncu --set full --clock-control none -k regex:'lda_tile|transpose' -s 3 -c 1 -f -o $O/r02_ncu_pack_copy tools/gemm_probe pack 8192 > $O/r02_ncu_pack_copy.log 2>&1
      timeout 600 ncu --set full --clock-control none -k regex:'lda_tile|transpose' -s 16 -c 1 -f -o $O/r02_ncu_pack_axpby tools/gemm_probe pack 8192 > $O/r02_ncu_pack_axpby.log 2>&1
      timeout 600 ncu --set full --clock-control none -k regex:'lda_tile|transpose' -s 29 -c 1 -f -o $O/r02_ncu_pack_transpose tools/gemm_probe pack 8192 > $O/r02_ncu_pack_transpose.log 2>&1
      for k in copy axpby transpose; do python tools/ncu_summary.py $O/r02_ncu_pack_$k.ncu-rep $O/r02_ncu_pack_$k.csv > /dev/null 2>&1; tail -n +3 $O/r02_ncu_pack_$k.csv | cut -c1-300; done
      ;;
    pending1)
      timeout 600 python tools/bench_configs.py --pending > $O/r02_bench_pending_1gpu.jsonl 2> $O/r02_bench_pending_1gpu.err
      cut -c1-400 $O/r02_bench_pending_1gpu.jsonl; tail -3 $O/r02_bench_pending_1gpu.err
      ;;
    *) echo "unknown section $section";;
  esac
done
echo "=== done ($(date +%T))"
