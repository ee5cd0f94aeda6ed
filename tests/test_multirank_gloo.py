"""N>1 path of the host layer on CPU: world_size 2 and 3 gloo jobs (one process per "GPU") over the oracle-bound
backend table.  Checks the factory's partitioning, the q-vector / atom / moment sharding inside a partition, the
all-reduce of packed partials and that exactly the partition-rank-0 processes write results."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np
import pytest

from sassena_b200 import host, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, case, tmp_path, extra=()):
    out = str(tmp_path / f"{case}_{world}.pkl")
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_gloo_worker.py"), out, case, *extra], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = []
    for p in procs:
        o, _ = p.communicate(timeout=120)
        logs.append(o.decode()[-2000:])
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    return pickle.load(open(out, "rb"))


def _reference(oracle, case):
    NA, NF = 23, 12
    xyz = synth.trajectory(NF, NA, 20.0, 0.2, 3, offset=-10.0)
    b = synth.factors(NA)
    qv = host.create_from_scans([{"base": (1, 0, 0), "from": 0.3, "to": 1.5, "points": 5}])
    out = []
    for q in qv:
        ql = np.linalg.norm(q)
        if case.startswith("all"):
            out.append(oracle.compute_all_vectors(xyz, b, ql * synth.unit_vectors(7, 1)))
        elif case.startswith("scan"):
            out.append(oracle.compute_all_vectors(xyz, b, ql * synth.unit_vectors(9, 1)))
        elif case.startswith("cyl"):
            out.append(oracle.compute_mpcylinder(oracle.cart_to_cylindrical(xyz, (1, 0, 1)), b, q, (1, 0, 1),
                                                 oracle.moments_cylinder(2)))
        elif case.startswith("self"):
            out.append(oracle.compute_self_vectors(xyz.transpose(1, 0, 2), b, ql * synth.unit_vectors(3, 1)))
        else:
            out.append(oracle.compute_mpsphere(oracle.cart_to_spherical(xyz), b, ql, oracle.moments_sphere(3), dsp="square"))
    return qv, out


@pytest.mark.parametrize("world,case", [(2, "all"), (2, "self"), (2, "mp"), (2, "all_manual1"), (3, "self"),
                                        (3, "all_manual1"), (2, "all_frames"), (3, "all_frames"), (2, "scan"), (2, "scan_frames"),
                                        (3, "scan_frames"), (2, "self_stream"), (3, "self_stream"), (2, "cyl"), (3, "cyl")])
def test_multirank_matches_single_rank(oracle, tmp_path, world, case):
    gathered = _run(world, case, tmp_path)
    qv, ref = _reference(oracle, case)
    records = {}
    writers = 0
    for rank, has, recs, timer_keys in gathered:
        assert has, "no spare ranks expected in these cases"
        assert "sd:compute" in timer_keys and "sd:stage" in timer_keys
        assert ("sd:c:b:exchange" in timer_keys) == ("_frames" in case)  # amplitude exchange only when frame-sharded
        assert ("sd:c:scan" in timer_keys) == case.startswith("scan")  # |q| batching needs >= 8 subvectors
        writers += 1 if recs else 0
        for r in recs:
            key = tuple(np.round(r["q"], 12))
            assert key not in records, "a |q| was written twice"
            records[key] = r
    if case.endswith("_manual1"):
        assert writers == min(world, len(qv))  # every single-rank partition writes its own |q| subset
    else:
        assert writers == 1  # one partition: only its rank 0 writes (abstract_scatter_device.cpp:239-244)
    assert len(records) == len(qv)
    for q, (rfqt, rfq, rfq2) in zip(qv, ref):
        r = records[tuple(np.round(q, 12))]
        scale = abs(rfqt[0])
        assert np.max(np.abs(r["fqt"] - rfqt)) < 1e-11 * scale
        assert abs(r["fq"] - rfq) < 1e-11 * scale
        assert abs(r["fq2"] - rfq2) < 1e-11 * abs(rfq2)


def test_replicated_staging_is_budgeted_up_front(oracle, tmp_path):
    """multipole devices keep every frame on every rank of a partition, so a partition of two does not halve a rank's share of
    the coordinates: with limits.stage.memory.data at 3/4 of the trajectory the plan must refuse up front (it used to accept
    NNPP = 2 on the frame-split byte count and fail in the stager after the trajectory had been read); the coherent device,
    which does split the frames, runs under the same budget"""
    for rank, has, recs, _ in _run(2, "mp_tight", tmp_path):
        assert not has and "Automatic decomposition failed" in recs
    gathered = _run(2, "all_frames_tight", tmp_path)
    assert all(has for _, has, _, _ in gathered) and sum(len(recs) for _, _, recs, _ in gathered) == 5


@pytest.mark.parametrize("case", ["all_manual2", "self_manual2"])
def test_spare_rank_does_not_hang_the_partition_split(oracle, tmp_path, case):
    """three ranks, manual partitions of two: the plan uses ranks 0-1 and leaves rank 2 spare (the reference allows this,
    scatter_device_factory.cpp:104-116).  The spare rank returns without a device and the others must not wait for it in the
    second communicator split."""
    gathered = _run(3, case, tmp_path)
    qv, ref = _reference(oracle, case)
    assert sorted(has for _, has, _, _ in gathered) == [False, True, True]
    records = {}
    for rank, has, recs, _ in gathered:
        assert has or not recs
        for r in recs:
            records[tuple(np.round(r["q"], 12))] = r
    assert len(records) == len(qv)
    for q, (rfqt, rfq, rfq2) in zip(qv, ref):
        r = records[tuple(np.round(q, 12))]
        assert np.max(np.abs(r["fqt"] - rfqt)) < 1e-11 * abs(rfqt[0]) and abs(r["fq"] - rfq) < 1e-11 * abs(rfqt[0])


@pytest.mark.parametrize("manual", [False, True])
def test_multirank_job_from_config(oracle, tmp_path, manual):
    """scatter.xml -> Job.run on 2 ranks: one partition of 2 (rank 0 writes all |q|) and, with manual partitions of
    size 1, two writers; load_signal merges the per-rank rows and they match the oracle."""
    from test_control_plane import ORIENT, SCAN, make_case
    limits = ""
    if manual:
        limits = ("<limits><decomposition><utilization>0</utilization><partitions><automatic>false</automatic>"
                  "<size>1</size></partitions></decomposition></limits>")
    cfg, xyz, names = make_case(tmp_path, scattering=SCAN + ORIENT, stager=limits)
    sig_dir = tmp_path / "signal"
    gathered = _run(2, "job", tmp_path, extra=(cfg, str(sig_dir)))
    written = sorted(w for _, w, _ in gathered)
    assert written == ([1, 2] if manual else [0, 3])
    sig = host.load_signal(sig_dir)
    job = host.Job(cfg)
    qv = job.qvectors()
    assert sorted(map(tuple, sig["qvectors"])) == sorted(map(tuple, qv))
    p = job.params()
    for i, q in enumerate(sig["qvectors"]):
        fqt, fq, fq2 = oracle.compute_all_vectors(xyz, job.factors(np.linalg.norm(q)), p.init_subvectors(q))
        assert np.allclose(sig["fqt"][i], fqt, rtol=1e-11, atol=1e-11 * abs(fqt[0]))
        assert np.isclose(sig["fq"][i], fq, rtol=1e-11) and np.isclose(sig["fq2"][i], fq2, rtol=1e-11)


@pytest.mark.parametrize("world,kind", [(2, "all"), (3, "all"), (2, "self"), (3, "self")])
def test_multirank_stager_dump(oracle, tmp_path, world, kind):
    """stager.dump on several ranks: every rank of the partition writes its own frames (DivAssignment, coherent with the
    reference's frame decomposition) or atom timelines (ModAssignment, self) into one DCD file (data_stager.cpp:131-165,352-391)"""
    from test_control_plane import make_case
    stager = "<stager><dump>true</dump><file>staged.dcd</file></stager>"
    # uneven blocks on purpose (9 frames / 11 atoms over 2 or 3 ranks): the utilization threshold is lowered for it
    stager += ("<limits><decomposition><utilization>0.5</utilization>"
               + ("<coherent>frames</coherent>" if kind == "all" else "") + "</decomposition></limits>")
    cfg, xyz, names = make_case(
        tmp_path, NA=11, NF=9, stager=stager,
        scattering=f"<type>{kind}</type><vectors><type>single</type><single><x>0.5</x><y>0</y><z>0</z></single></vectors>"
                   "<average><orientation><type>none</type></orientation></average>")
    gathered = _run(world, "job", tmp_path, extra=(cfg, str(tmp_path / "signal")))
    assert sorted(w for _, w, _ in gathered) == [0] * (world - 1) + [1]
    got = host.DCDFile(str(tmp_path / "staged.dcd")).read()
    if kind == "all":
        assert np.array_equal(got, xyz)
    else:
        assert np.array_equal(got, xyz.transpose(1, 0, 2))


@pytest.mark.parametrize("world,mode", [(2, "frames"), (3, "atoms")])
def test_multirank_stage_only(oracle, tmp_path, world, mode):
    """the s_stage flow (s_stage.cpp:205-232) on several ranks: every rank stages and dumps its DivAssignment block of frames
    (stager.mode = frames) or its ModAssignment atoms (atoms); the bytes staged add up to the whole trajectory"""
    from test_control_plane import SCAN, make_case
    cfg, xyz, names = make_case(tmp_path, NA=11, NF=9, scattering=SCAN,
                                stager=f"<stager><dump>true</dump><file>staged.dcd</file><mode>{mode}</mode></stager>")
    gathered = _run(world, "stage", tmp_path, extra=(cfg, str(tmp_path / "unused")))
    assert sum(n for _, n, _ in gathered) == xyz.size * 4
    assert all(f"stager.mode={mode}" in rep for _, _, rep in gathered)
    got = host.DCDFile(str(tmp_path / "staged.dcd")).read()
    assert np.array_equal(got, xyz if mode == "frames" else xyz.transpose(1, 0, 2))
    assert not os.path.exists(tmp_path / "unused")
