"""bench.py's output contract: exactly one JSON line on stdout carrying the keys the driver reads.  The reference arm
runs on the CPU and is checked everywhere; the CUDA arm is checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ)
    env.pop("RANK", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "1", "--workload", "c1"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "fp64_sincos_throughput" and j["unit"] == "Gelem/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["dtype"] == "f64" and j["data"] == "synthetic" and j["vs_baseline"] is None and j["value"] > 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    # the workload part of `config` is the CUDA arm's, key for key: the two arms are on the same configuration
    sys.path.insert(0, ROOT)
    import bench

    assert j["config"] == bench.workload_config("c1", 1_000_000) and j["setup"]["elements_per_step"] == 1_000_000


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""



@pytest.mark.gpu
def test_cuda_arm_prints_one_json_line_with_the_contract_keys():
    """The arm the driver runs on the B200: one JSON line with value / roofline / e2e / cpu_baseline / clocks /
    gpu_launches (on a reduced element count so that the test takes seconds)."""
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "4", "--warmup", "3",
                          "--elems", "80000000", "--e2e-elems", "20000000", "--pageable-elems", "10000000", "--sustain-seconds", "0.3"],
                         capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    j = json.loads(lines[0])
    assert j["metric"] == "fp64_sincos_throughput" and j["unit"] == "Gelem/s" and j["dtype"] == "f64" and j["n_gpus"] == 1
    assert j["steps"] == 4 and j["warmup"] == 3 and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["value"] > 50 and j["gpu_launches"] >= 4
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["algorithmic_bytes_per_launch"] == 24 * 80000000 and "traffic" in r and r["traffic_source"]
    assert r["traffic"] is None or 0.9 * 24 * 80000000 < r["traffic"] < 1.2 * 24 * 80000000
    sus = r["sustained"]
    assert sus["seconds"] >= 0.25 and sus["steps"] >= 10 and 0.3 < sus["frac"] < 1.2 and "sm_mhz" in sus["clocks"]
    e = j["e2e"]
    assert e["value"] > 0.5 and e["h2d_bytes_per_step"] == 8 * 20000000 and e["d2h_bytes_per_step"] == 16 * 20000000
    er = e["roofline"]   # the host link's own ceiling, measured in the run by a bare-copy probe on the same buffers
    assert er["bound"] in ("pcie", "host_dram") and er["peak_gbs"] > 20 and 0.3 < er["frac"] < 1.15
    assert abs(er["frac"] - e["value"] / er["peak"]) < 1e-9
    ha = e["with_host_assist"]   # reported beside e2e, never instead of it: the same call with option host_assist
    assert ha["value"] > 0.5 and 0.0 < ha["gpu_share"] <= 1.0 and ha["host_threads_per_rank"] >= 1 and "NOT the headline" in ha["note"]
    pg = j["e2e_pageable"]
    assert pg["value"] > 0.2 and pg["elements_per_gpu"] == 10000000 and "void fabe13_sincos" in pg["path"]
    assert j["config"]["elements_total"] == 80000000 and j["setup"]["elements_per_gpu"] == 80000000
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["value"] > 0 and j["cpu_baseline"]["cores"] >= 1
    assert "sm_mhz" in j["clocks"] and "reasons" in j["clocks"] and "workload" in j["config"]


def test_ncu_traffic_is_carried_over_only_for_the_kernel_sources_it_was_captured_on(tmp_path, monkeypatch):
    """roofline.traffic comes from the committed ncu capture (profiles/ncu_traffic.json).  bench.py uses it only while the
    capture's stamp equals the hash of the kernel sources of this build; otherwise it prints null and says why."""
    sys.path.insert(0, ROOT)
    import bench

    now = bench.kernel_source_hash()
    assert len(now) == 64
    fake_root = tmp_path
    (fake_root / "profiles").mkdir()
    (fake_root / "fabe_b200" / "csrc").mkdir(parents=True)
    for name in bench.KERNEL_SOURCES:
        (fake_root / "fabe_b200" / "csrc" / name).write_bytes(open(os.path.join(ROOT, "fabe_b200", "csrc", name), "rb").read())
    monkeypatch.setattr(bench, "ROOT", str(fake_root))
    assert bench.kernel_source_hash() == now
    per, why = bench.ncu_traffic_per_elem()
    assert per is None and "no ncu capture" in why
    stamp = {"dram_bytes_per_elem": 23.7, "report": "x.ncu-rep", "kernel_source_sha256": now}
    (fake_root / "profiles" / "ncu_traffic.json").write_text(json.dumps(stamp))
    per, why = bench.ncu_traffic_per_elem()
    assert per == 23.7 and now[:12] in why
    with open(fake_root / "fabe_b200" / "csrc" / bench.KERNEL_SOURCES[0], "ab") as f:
        f.write(b"\n// edited after the capture\n")
    per, why = bench.ncu_traffic_per_elem()
    assert per is None and "not carried over" in why
