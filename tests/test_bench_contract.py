"""bench.py contract on the CPU: the reference arm (`--impl reference`, the oracle port on the host cores) prints exactly
one JSON line with the keys the driver reads, for the coherent and the self workload; helper arithmetic matches the
host layer's decomposition."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEYS = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches")


@pytest.mark.parametrize("workload", ["C3", "C2", "C4"])
def test_reference_arm_prints_one_json_line(workload):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--frames", "64", "--atoms", "300", "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.02"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "evals/s"
    # coherent: the reference's own AllVectorsScatterDevice when oracle/_ref is built, else the oracle port; self / multipole: port
    from oracle import oracle as o
    want = "reference" if (workload == "C3" and o.have_ref_smath()) else "port"
    assert d["cpu_baseline"]["kind"] == want and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_partition_helpers_match_host_layer():
    import bench
    from sassena_b200 import host
    for NN, N in ((1, 7), (2, 7), (3, 10), (8, 30000), (8, 5)):
        for r in range(NN):
            assert bench.div_assignment(NN, r, N) == host.div_assignment(NN, r, N)[:2]
            assert bench.mod_assignment_count(NN, r, N) == host.mod_assignment(NN, r, N)[1]
    # pass sizes of the scan kernel cover the batch exactly, none longer than the kernel's largest pass
    for nq in (1, 4, 27, 28, 29, 50, 200):
        assert sum(bench.scan_passes(nq)) == nq and max(bench.scan_passes(nq)) <= bench.SCAN_MAX_PASS
        assert sum(bench.scan_passes(nq, bench.SCAN_MAX_PASS_CORR)) == nq
    assert bench.scan_passes(50) == [25, 25] and bench.scan_passes(50, bench.SCAN_MAX_PASS_CORR) == [17, 17, 16]
    flop, instr = bench.scan_work_per_eval(50, False)
    assert 4.0 < instr < 5.0 and flop > instr  # 38 set-up + 12 pairs x 6 per 25 |q|


def test_cpu_arm_uses_every_core_not_omp_num_threads():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use all cores this process may run on"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "all", "--skip", "C5s",
                          "--frames", "64", "--atoms", "300", "--steps", "1", "--warmup", "1", "--cpu-seconds", "0.02"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    cores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == cores
    assert set(d["workloads"]) == {"C2", "C4"} and all(w["cpu_baseline"]["cores"] == cores for w in d["workloads"].values())


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [["--workload", "C3", "--atoms", "3000", "--frames", "300"],
                                   ["--workload", "C3", "--atoms", "3000", "--frames", "300", "--mode", "scan"],
                                   ["--workload", "C3", "--atoms", "3000", "--frames", "300", "--mode", "per-q"],
                                   ["--workload", "C2", "--atoms", "96", "--frames", "9000"],
                                   ["--workload", "C4", "--atoms", "5000", "--frames", "40"],
                                   ["--workload", "C5s", "--atoms", "96", "--frames", "9000", "--wave-atoms", "40"]])
def test_bench_ours_small_sizes_json_line(extra):
    """the product arm on a reduced workload (debug overrides): one JSON line with roofline, cpu_baseline, e2e, clocks,
    launches > 0 and parity against the oracle inside the tolerance"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--cpu-seconds", "0.05"]
                         + extra, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["n_gpus"] == 1
    assert d["roofline"]["bound"] == "fp64" and 0 < d["roofline"]["frac"] < 1.5 and d["roofline"]["peak"] > 10
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    assert d["parity"]["fqt_rel_err"] < 1e-9 and d["parity"]["fq_rel_err"] < 1e-9
    assert "workload" in d["config"]
    if "C5s" in extra:  # streamed: three waves, same result as the resident run (up to the order of the sums over atoms)
        assert d["config"]["waves_per_rank"] == 3 and d["staging"]["resident"]["streamed_vs_resident_rel_err"] < 1e-11


@pytest.mark.gpu
def test_bench_default_invocation_carries_every_workload():
    """the default invocation (reduced sizes here): the headline line on the reference's own |q| generator (corrected scan
    kernel planned) plus the equally spaced scan, C2, C4 and the streamed C5 sample under `workloads`"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--cpu-seconds", "0.05",
                          "--atoms", "2000", "--frames", "400", "--c5s-atoms", "64"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["config"]["mode"] == "scan-rounded" and d["config"]["scan_plan_plain_corrected_single"][1] > 0
    assert set(d["workloads"]) == {"C3_equally_spaced", "C2", "C4", "C5s"}
    for name, w in d["workloads"].items():
        assert "error" not in w, (name, w)
        assert w["value"] > 0 and w["gpu_launches"] > 0 and w["roofline"]["frac"] > 0 and w["e2e"]["value"] > 0, name
        if name != "C3_equally_spaced":
            assert w["parity"]["fqt_rel_err"] < 1e-9 and w["cpu_baseline"]["value"] > 0, name
    assert d["workloads"]["C3_equally_spaced"]["config"]["scan_plan_plain_corrected_single"][0] > 0
