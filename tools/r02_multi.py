#!/usr/bin/env python
"""Round-2 multi-GPU sessions (one gpurun call on N = 4 or 8 GPUs; every minute is charged N times, so the order is: the
measurement that decides the defaults first, parity of the winner second, everything else after).

  gpurun --gpus 4 --timeout 1000 -- 'python tools/r02_multi.py'                                   (session 3: knob A/B, 46 GPU-min)
  gpurun --gpus 4 --timeout 900  -- 'python tools/r02_multi.py bench e2e parity configs'          (session 5: shipped defaults, 33)
  gpurun --gpus 8 --timeout 800  -- 'python tools/r02_multi.py bench e2e parity dropin configs'   (session 6: 86 — `parity`
                                                                                                   includes the pending group: 4 min x 8)
Stages (all by default, or the ones named on the command line):
  bench    bench.py, device-resident leg with rank 0's per-launch timeline, at the library defaults and with each candidate
           set of knobs (session 3: merge-panels / panel-transport / bg-ctas combinations; later: fused-reduce 2 on 8 GPUs);
           the fastest valid line wins and its knobs are handed to the later stages
  parity   tests/dist_worker.py (oracle + the reference's golden outputs), then its pending group (widening rows)
  dropin   the reference's own unmodified test mains under tools/candmc_run (tests/test_dropin_gpu.py), its QR tests over the
           upd_A / cdgemm seams and the cases no B200 has run yet (tests/test_zz_{qr_dropin,unseen,aggregator}_gpu.py)
  e2e      bench.py with the end-to-end leg (both passes) and timelines of the end-to-end steps
  configs  tools/bench_configs.py: BASELINE configs 2 / 4 / 5 with result checks
Everything goes to gpurun_out/r02_*_{N}gpus*; stdout carries a summary.  R02_SIM=<ranks> dry-runs the script's own logic on
the CPU simulator (numbers of such a run mean nothing)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
T0 = time.time()


SIM = os.environ.get("R02_SIM")   # dry run of this script's own logic on the CPU simulator: R02_SIM=<ranks>


def ngpus():
    if SIM:
        return int(SIM)
    import torch

    return torch.cuda.device_count()


NG = ngpus()
PORT = [29600]


def run(tag, script_args, env=None, timeout=300, torchrun=True):
    """one torchrun job (its own process group, killed as a group on timeout); returns (rc, stdout)"""
    PORT[0] += 1
    cmd = ([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={NG}", "--master-addr", "127.0.0.1",
            "--master-port", str(PORT[0])] if torchrun else [sys.executable]) + script_args
    e = dict(os.environ)
    e.update(env or {})
    e.setdefault("NCCL_DEBUG", "WARN")
    if SIM:
        e.update(CANDMC_CPUSIM="1", CPUSIM_TIMEOUT="60", OMP_NUM_THREADS="1")
        if script_args[0].endswith(".py") and "dist_worker" not in script_args[0]:
            e["CPUSIM_ARGS"] = "--n 512" if script_args[0] == "bench.py" else "--shrink 64"   # (torchrun's parser trips over --n)
            cmd = cmd[:len(cmd) - len(script_args)] + ["tests/cpusim/run_sim.py"] + script_args
    t = time.time()
    with open(os.path.join(OUT, tag + ".out"), "w") as so, open(os.path.join(OUT, tag + ".err"), "w") as se:
        p = subprocess.Popen(cmd, cwd=ROOT, env=e, stdout=so, stderr=se, start_new_session=True)
        try:
            rc = p.wait(timeout=timeout)
        except subprocess.TimeoutExpired:
            import signal

            os.killpg(p.pid, signal.SIGKILL)
            p.wait()
            rc = -9
    txt = open(os.path.join(OUT, tag + ".out")).read()
    print(f"[{time.time() - T0:6.0f}s] {tag}: rc={rc} ({time.time() - t:.0f}s)", flush=True)
    return rc, txt


def last_json(txt):
    for line in reversed(txt.splitlines()):
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                pass
    return None


def knob_env(knobs):
    """the same switches for scripts that cannot take bench.py's flags: the library reads them when it binds to its GPU"""
    env = {}
    if "--merge-panels" in knobs:
        env["CANDMC_MERGE_PANELS"] = knobs[knobs.index("--merge-panels") + 1]
        env["CANDMC_TEST_MERGE_PANELS"] = env["CANDMC_MERGE_PANELS"]
    if "--panel-transport" in knobs:
        env["CANDMC_PANEL_TRANSPORT"] = "1"
        env["CANDMC_TEST_PANEL_TRANSPORT"] = "1"
    if "--no-panel-transport" in knobs:
        env["CANDMC_PANEL_TRANSPORT"] = "0"
    if "--fused-reduce" in knobs:
        env["CANDMC_FUSED_REDUCE"] = knobs[knobs.index("--fused-reduce") + 1]
        if env["CANDMC_FUSED_REDUCE"] == "2":
            env["CANDMC_TEST_FUSED_GRIDS"] = "1"
    if "--bg-ctas" in knobs:
        env["CANDMC_BG_CTAS"] = knobs[knobs.index("--bg-ctas") + 1]
    return env


def main():
    only = sys.argv[1:]   # optional: names of the stages to run (bench parity dropin e2e configs)
    want = lambda s: not only or s in only  # noqa: E731
    summary = {"n_gpus": NG}
    best_knobs = []
    if want("bench"):
        # the library defaults (merged launches + copy-engine panels since the first 4-GPU session of this round) and what is
        # still opt-in: the fused depth sum on q x q x c grids; the round-1 schedule once for the record
        cands = [[]]
        if NG == 8:
            cands += [["--fused-reduce", "2"]]
        if os.environ.get("R02_WITH_ROUND1"):
            cands += [["--merge-panels", "0", "--no-panel-transport"]]
        best = None
        for kn in cands:
            tag = f"r02_bench{NG}_" + ("default" if not kn else "_".join(x.strip("-").replace("-", "") for x in kn))
            rc, txt = run(tag, ["bench.py", "--gpus", str(NG), "--steps", "4", "--warmup", "3", "--no-e2e", "--timeline"] + kn, timeout=150)
            line = last_json(txt)
            rel = line.get("rel_frobenius_vs_cublas_crosscheck") if line else None
            ok = rc == 0 and rel is not None and rel <= line.get("tolerance_10_n_eps", 0)
            rec = {"knobs": kn, "rc": rc, "valid": bool(ok)}
            if line:
                rec.update(value=line["value"], ms=line["ms_per_step"], exposed=line.get("exposed_non_gemm_pct"),
                           avg_launch_ms=line["roofline"]["avg_launch_ms"], launches=line["roofline"]["launches"],
                           rel=line.get("rel_frobenius_vs_cublas_crosscheck"))
            print("   ", json.dumps(rec), flush=True)
            summary.setdefault("bench", []).append(rec)
            if ok and (best is None or line["value"] > best[0]):
                best = (line["value"], kn)
        best_knobs = best[1] if best else []
        summary["winner"] = best_knobs
        print("winner:", best_knobs, flush=True)
    elif os.environ.get("R02_KNOBS"):
        best_knobs = os.environ["R02_KNOBS"].split()
    kenv = knob_env(best_knobs)
    if want("parity"):
        rc, txt = run(f"r02_parity_{NG}gpus", ["tests/dist_worker.py"], env=dict(kenv, CANDMC_TEST_VERBOSE="1"), timeout=420)
        summary["parity"] = {"rc": rc, "line": last_json(txt), "knobs": best_knobs}
        print("   ", json.dumps(summary["parity"])[:400], flush=True)
        rc, txt = run(f"r02_parity_pending_{NG}gpus", ["tests/dist_worker.py"], env=dict(kenv, CANDMC_TEST_PENDING="1", CANDMC_TEST_VERBOSE="1"),
                      timeout=360)
        summary["parity_pending"] = {"rc": rc, "line": last_json(txt)}
        print("   ", json.dumps(summary["parity_pending"])[:400], flush=True)
    if want("dropin"):
        rc, txt = run(f"r02_dropin_{NG}gpus", ["-m", "pytest", "tests/test_dropin_gpu.py", "tests/test_zz_qr_dropin_gpu.py", "tests/test_zz_unseen_gpu.py",
                                                 "tests/test_zz_aggregator_gpu.py", "-m", "gpu", "-q", "-rA", "-p", "no:cacheprovider"],
                      env=kenv, timeout=300, torchrun=False)
        summary["dropin"] = {"rc": rc, "tail": txt.strip().splitlines()[-1:] if txt.strip() else []}
        print("   ", json.dumps(summary["dropin"]), flush=True)
    if want("e2e"):
        rc, txt = run(f"r02_bench{NG}_e2e", ["bench.py", "--gpus", str(NG), "--steps", "4", "--warmup", "3", "--timeline"] + best_knobs, timeout=500)
        line = last_json(txt)
        summary["e2e"] = {"rc": rc, "value": line and line["value"], "e2e": line and line.get("e2e"),
                          "passes": line and [(p.get("host_operand_settings"), p.get("value"), p.get("valid")) for p in line.get("e2e_passes", [])]}
        print("   ", json.dumps(summary["e2e"])[:900], flush=True)
    if want("configs"):
        rc, txt = run(f"r02_configs_{NG}gpus", ["tools/bench_configs.py"], env=kenv, timeout=400)
        summary["configs"] = [json.loads(l) for l in txt.splitlines() if l.startswith("{")]
        for l in summary["configs"]:
            print("   ", json.dumps({k: l.get(k) for k in ("config", "ms", "tflops", "pct_of_fp64_tensor_peak", "check_passed",
                                                           "rel_frobenius_vs_cublas_crosscheck")}), flush=True)
    with open(os.path.join(OUT, f"r02_multi_summary_{NG}gpus.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(f"done in {time.time() - T0:.0f}s", flush=True)


if __name__ == "__main__":
    main()
