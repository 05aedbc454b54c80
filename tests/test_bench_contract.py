"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the oracle port on the host
cores and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line():
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["unit"] == "state-time updates/s" and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # same `config` as the native arm prints for the same K, T, N (the bounded sample is described in cpu_baseline)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == json.loads(json.dumps(bench.workload_config(100000, 1000, 1)))
    assert d["steps"] == 1 and d["warmup"] == 0
    lr = cb["live_reference"]
    assert lr is None or (lr["extrapolated"]["label"].startswith("EXTRAPOLATED") and len(lr["rows"]) == 3)


def test_parity_block_helpers_pick_the_wrap_around_state():
    """The columns bench.py's parity block checks always include the first hidden state and the wrap-around state,
    and the oracle names exactly the rows that have to be gathered for them."""
    from oracle import vgm_oracle as vo
    K = 100000
    hidden = list(range(1, K, 2))
    need_s, need_l = vo.lorenz96_column_needs([1, 4711, K - 1], K, hidden)
    assert need_l == [1] and set([K - 4, K - 3, K - 2, K - 1, 0, 1, 2]) <= set(need_s) and len(need_s) <= 21


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, cwd=ROOT,
                         timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""
