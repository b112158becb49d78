"""The bench.py JSON contract (driver-facing): keys, units and the reference arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--db-proteins", "3000", "--queries", "800", "--ref-queries", "400", "--batches", "2",
         "--c4-proteins", "30000", "--c4-sample", "6", "--sustain-s", "0.2"]


def _run(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_cpu():
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"] + SMALL)
    assert d["impl"] == "reference" and d["metric"] == "query residues/sec" and d["unit"] == "residues/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"] + SMALL,
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_bench_line_gpu():
    d = _run(["--steps", "3", "--warmup", "3"] + SMALL)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "query residues/sec" and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["sustained"]["value"] > 0 and d["sustained"]["steps"] >= d["steps"]
    assert "go" in d["reference_toolchain"]
    # the C4 block (here at toy size): device-generated DB, streaming build, parity sample against the oracle
    c4 = d["c4"]
    assert "error" not in c4, c4
    assert c4["value"] > 0 and c4["status_flags"] == 0 and c4["roofline"]["frac"] > 0
    assert c4["parity_sample"]["mismatches"] == 0 and c4["parity_sample"]["kstats_equal"] is True
    st = d["stages"]
    assert "error" not in st["c5"] and st["c5"]["parity_spot_check"]["dp_score_and_raw_equal_oracle"] is True
    assert "error" not in st["c2"] and st["c2"]["value"] > 0


def test_traffic_json_was_captured_from_the_current_kernel_sources():
    """bench.py reports roofline.traffic only while profiles/traffic.json carries the hash of the kernel sources its
    ncu captures were taken from; a kernel edit without a new capture must show up here, not as a silent null"""
    import json

    sys.path.insert(0, ROOT)
    import bench

    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert {"k_search_wt<W>", "k_search_m", "k_search_f<CLS4>", "k_search_f<CLS5>", "k_search_f<CLS6>"} <= set(t["kernels"])
    for name, e in t["kernels"].items():
        assert e["source_sha16"] == bench.kernel_source_hash(), name
        assert e["dram_bytes_per_launch"] > 0
    assert bench.traffic_from_profiles("k_search_wt<W>")[0] > 1e9
    assert bench.c4_traffic()["traffic"] > 5e9
