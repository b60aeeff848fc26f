"""bench.py's reference arm and config block (CPU only: the repo arm needs a B200)."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_config_block_is_static_and_shared_by_both_arms():
    """`config` holds only static facts of the workload, so `--impl reference` prints the same dictionary as the repo arm."""
    bench = _bench()
    for name, wl in bench.WORKLOADS.items():
        args = types.SimpleNamespace(workload=name, inputs="default", global_batch=0)
        a, b = bench.workload_config(args, wl), bench.workload_config(args, wl)
        assert a == b and a["workload"] == name and "l2" in a
        json.dumps(a)
    # strong scaling: rank 0's shard of the global batch, identical in both arms
    args = types.SimpleNamespace(workload=bench.DEFAULT_WORKLOAD, inputs="default", global_batch=64)
    os.environ["WORLD_SIZE"] = "8"
    try:
        cfg = bench.workload_config(args, bench.WORKLOADS[bench.DEFAULT_WORKLOAD])
    finally:
        del os.environ["WORLD_SIZE"]
    assert cfg["B_per_gpu"] == 8 and cfg["global_batch"] == 64


def test_reference_arm_line_on_the_ci_workload():
    """`bench.py --impl reference` (CPU): one JSON line with the contract's keys; `config` equals the repo arm's block"""
    bench = _bench()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny_256x512_b2_n12",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "decoded Mpix/s" and line["unit"] == "Mpix/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    args = types.SimpleNamespace(workload="tiny_256x512_b2_n12", inputs="default", global_batch=0)
    assert line["config"] == bench.workload_config(args, bench.WORKLOADS["tiny_256x512_b2_n12"])


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload",
                          "tiny_256x512_b2_n12"], capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
