"""Multi-GPU parity check, run under torchrun (one process per GPU, NCCL): the C++ host layer's factory + devices
with a torch.distributed communicator against the CPU oracle.  Usage:
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
    for name, prm, ref_fn in cases:
        recs, has, tm = host.run_scatter(prm, xyz, qv, b=b, comm=comm)
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
            print(f"[{world} GPUs] {name:24s} writers={sum(1 for rr in allrecs if rr)} max rel err vs oracle = {err:.2e}")
    if rank == 0:
        assert worst < 1e-9, worst
        print("multigpu_check OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
