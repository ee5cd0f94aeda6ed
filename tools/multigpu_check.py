"""Multi-GPU parity check, run under torchrun (one process per GPU, NCCL): the C++ host layer's factory + devices against the
CPU oracle, once with the library's own NCCL communicator (native callbacks; the frame-sharded coherent device then exchanges
amplitudes inside the library) and once with a torch.distributed communicator (Python callbacks; amplitude all-reduce), plus
the C-ABI's sharded scan called directly.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multigpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import oracle as o  # noqa: E402
from sassena_b200 import host, synth  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = host.TorchDistCommunicator(None, device_memory=True)
    NA, NF = 400, 64
    xyz = synth.trajectory(NF, NA, 30.0, 0.2, 41, offset=-15.0)
    b = synth.factors(NA)
    qv = host.create_from_scans([{"base": (1, 0, 0), "from": 0.2, "to": 2.0, "points": 5}])
    worst = 0.0
    cases = []
    p = host.Params().set("scattering.average.orientation.type", "vectors")
    p.set("scattering.average.orientation.vectors.resolution", 37).create()
    p.set("limits.decomposition.coherent", "frames")
    cases.append(("all (frame-sharded)", p, lambda q, p=p: o.compute_all_vectors(xyz, b, p.init_subvectors(q), nthreads=4)))
    pv = host.Params().set("scattering.average.orientation.type", "vectors")
    pv.set("scattering.average.orientation.vectors.resolution", 37).set("limits.decomposition.coherent", "vectors").create()
    cases.append(("all (vector-sharded)", pv, lambda q, p=pv: o.compute_all_vectors(xyz, b, p.init_subvectors(q), nthreads=4)))
    pq = host.Params().set("scattering.average.orientation.type", "vectors").set("scattering.dsp.type", "square")
    pq.set("scattering.average.orientation.vectors.resolution", 21).set("limits.decomposition.coherent", "frames").create()
    cases.append(("all square (frame-sh.)", pq,
                  lambda q, p=pq: o.compute_all_vectors(xyz, b, p.init_subvectors(q), dsp="square", nthreads=4)))
    ps = host.Params().set("scattering.type", "self").set("scattering.average.orientation.type", "vectors")
    ps.set("scattering.average.orientation.vectors.resolution", 5).create()
    cases.append(("self", ps, lambda q, p=ps: o.compute_self_vectors(xyz.transpose(1, 0, 2), b, p.init_subvectors(q), nthreads=4)))
    pm = host.Params().set("scattering.average.orientation.type", "multipole")
    pm.set("scattering.average.orientation.multipole.moments.type", "resolution")
    pm.set("scattering.average.orientation.multipole.moments.resolution", 6).create()
    cases.append(("mpsphere", pm, lambda q, p=pm: o.compute_mpsphere(o.cart_to_spherical(xyz), b, np.linalg.norm(q), p.moments, nthreads=4)))
    pp = host.Params().set("scattering.average.orientation.type", "vectors")
    pp.set("scattering.average.orientation.vectors.resolution", 11).create()
    pp.set("limits.decomposition.partitions.automatic", False).set("limits.decomposition.partitions.size", 1)
    pp.set("limits.decomposition.utilization", 0.0)
    cases.append(("all, 1-GPU partitions", pp, lambda q, p=pp: o.compute_all_vectors(xyz, b, p.init_subvectors(q), nthreads=4)))
    native = host.NcclCommunicator.from_torch_distributed(local)
    for cname, cm in (("native NCCL comm", native), ("torch.distributed comm", comm)):
        for name, prm, ref_fn in cases:
            recs, has, tm = host.run_scatter(prm, xyz, qv, b=b, comm=cm)
            allrecs = [None] * world
            dist.all_gather_object(allrecs, recs)
            if rank == 0:
                flat = [r for rr in allrecs for r in rr]
                assert len(flat) == len(qv), (name, len(flat))
                err = 0.0
                for r in flat:
                    ref = ref_fn(r["q"])
                    err = max(err, float(np.max(np.abs(r["fqt"] - ref[0])) / np.max(np.abs(ref[0]))))
                worst = max(worst, err)
                print(f"[{world} GPUs, {cname}] {name:24s} writers={sum(1 for rr in allrecs if rr)} max rel err vs oracle = {err:.2e}")
    native.close()

    # the C-ABI directly: frame-sharded scan with the exchange inside the library (float-rounded |q|: corrected kernel)
    import sassena_b200
    NA2, NF2, NM2 = 3000, 203, 61  # frame and subvector counts that do not divide evenly
    xyz2 = synth.trajectory(NF2, NA2, 60.0, 0.1, 7)
    b2 = synth.factors(NA2)
    u2 = synth.unit_vectors(NM2, 3)
    qls = synth.qlengths(0.1, 3.0, 19)
    ctx = sassena_b200.ScatterContext(local)
    box = [sassena_b200.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], world, rank)
    f_off, f_cnt = (rank * NF2) // world, ((rank + 1) * NF2) // world - (rank * NF2) // world
    ctx.stage_frames(np.ascontiguousarray(xyz2[f_off:f_off + f_cnt]))
    ctx.set_frame_window(NF2, f_off)
    ctx.set_factors(b2)
    plen = ctx.partial_len("autocorrelate")
    part = torch.zeros(len(qls) * plen, dtype=torch.float64, device="cuda")
    for rep in range(2):
        ctx.compute_all_vectors_scan_sharded(u2, qls, part.data_ptr())
    ctx.synchronize()
    got = [ctx.finalize(part.data_ptr() + n * plen * 8, 1.0 / NM2) for n in range(len(qls))]
    ctx.compute_all_vectors_sharded(qls[3] * u2, part.data_ptr())
    one = ctx.finalize(part.data_ptr(), 1.0 / NM2)
    if rank == 0:
        err = 0.0
        for n in (0, 3, len(qls) - 1):
            ref = o.compute_all_vectors(xyz2, b2, qls[n] * u2, nthreads=8)
            err = max(err, float(np.max(np.abs(got[n][0] - ref[0])) / np.max(np.abs(ref[0]))),
                      float(abs(got[n][1] - ref[1]) / abs(ref[0][0])))
            if n == 3:
                err = max(err, float(np.max(np.abs(one[0] - ref[0])) / np.max(np.abs(ref[0]))))
        worst = max(worst, err)
        print(f"[{world} GPUs] C-ABI sharded scan / single |q| (plan {ctx.last_scan_plan()}) max rel err vs oracle = {err:.2e}")
    ctx.close()
    if rank == 0:
        assert worst < 1e-9, worst
        print("multigpu_check OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
