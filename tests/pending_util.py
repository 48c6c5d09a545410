"""Helpers for the GPU tests of paths that have not run on a B200 yet (tests/test_zz_*.py).

Every case runs in its own process GROUP with a short timeout; on expiry the whole group is killed (no orphan keeps the
GPU), and once one case of a file has hung the remaining cases of that file are reported as xfail immediately, so that an
unexpected deadlock in new code costs minutes, not the session."""
import os
import signal
import subprocess

import pytest

_HUNG = {}


def run_guarded(key, cmd, timeout, cwd, env=None):
    """Returns (returncode, stdout, stderr).  `key` identifies the file whose later cases are skipped after a hang."""
    if _HUNG.get(key):
        pytest.xfail(f"an earlier case of {key} hung; not running the rest")
    p = subprocess.Popen(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        _HUNG[key] = True
        try:
            os.killpg(p.pid, signal.SIGKILL)
        except ProcessLookupError:
            pass
        out, err = p.communicate()
        pytest.fail(f"timed out after {timeout} s\n{out[-1500:]}\n{err[-1500:]}")
    return p.returncode, out, err
