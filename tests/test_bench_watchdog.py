"""bench.py's watchdog for its secondary blocks (`run_guarded`): a block that returns keeps its result and the timer
never fires; a block that does not return lets `on_timeout` print the line and end the process (exercised in a
subprocess, as bench.py does it with os._exit)."""
import importlib.util
import json
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_run_guarded_returns_the_result_and_cancels_the_timer():
    fired = []
    out = _bench().run_guarded("blk", 0.3, lambda: {"ok": 1}, lambda label, s: fired.append(label))
    import time

    time.sleep(0.5)
    assert out == {"ok": 1} and fired == []


def test_run_guarded_propagates_exceptions():
    import pytest

    with pytest.raises(ValueError):
        _bench().run_guarded("blk", 5, lambda: (_ for _ in ()).throw(ValueError("x")), lambda label, s: None)


def test_a_stuck_block_still_yields_the_line():
    code = textwrap.dedent(f"""
        import importlib.util, json, os, sys, time
        spec = importlib.util.spec_from_file_location("bench_module", r"{ROOT / 'bench.py'}")
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        state = {{"flux_c5": None}}
        def on_timeout(label, limit_s):
            state[label] = {{"error": f"watchdog: no result after {{limit_s}} s"}}
            print(json.dumps({{"value": 1.0, **state}}), flush=True)
            os._exit(0)
        mod.run_guarded("flux_c5", 0.5, lambda: time.sleep(60), on_timeout)
        print("not reached")
    """)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.0 and "watchdog" in line["flux_c5"]["error"]
