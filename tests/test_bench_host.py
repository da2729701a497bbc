"""CPU tests of bench.py's host logic (nothing here touches a GPU)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_DRIVER = textwrap.dedent("""
    import sys, time
    sys.path.insert(0, {root!r})
    import bench
    mode = sys.argv[1]
    strong = {{"ms_per_step": 1.0}}
    line = {{"metric": "m", "strong": strong}}
    def fn():
        if mode == "hang":
            time.sleep(60)
        if mode == "raise":
            raise RuntimeError("boom")
        return {{"ms_per_step": 0.5}}
    bench._finish_with_graph_strong(line, strong, fn, 0, limit_s=1.5)
    print("not reached")
""")


@pytest.mark.parametrize("mode", ["ok", "hang", "raise"])
def test_graph_strong_watchdog_prints_exactly_one_line(mode, tmp_path):
    """N > 1 bench lines end with the graph-captured sharded step measured under a watchdog: whatever that attempt does
    (returns, hangs in a collective, raises) the process prints ONE JSON line and exits 0."""
    drv = tmp_path / "drv.py"
    drv.write_text(_DRIVER.format(root=ROOT))
    r = subprocess.run([sys.executable, str(drv), mode], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    got = json.loads(lines[0])["strong"]
    assert got["ms_per_step"] == 1.0
    if mode == "ok":
        assert got["graph"] == {"ms_per_step": 0.5}
    else:
        assert "unavailable" in got["graph"]
