"""worker for test_multirank_gloo.py: one rank of a world_size-N gloo job driving the host layer over the
oracle-bound backend (CPU).  Writes the records gathered on rank 0 to the path in argv[1]."""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as dist  # noqa: E402

from oracle_backend import OracleBackend  # noqa: E402
from sassena_b200 import host, synth  # noqa: E402


def main():
    out_path, case = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    be = OracleBackend()
    comm = host.TorchDistCommunicator(None, device_memory=False)
    if case.startswith("job") or case == "stage":  # control plane: argv[3] = scatter.xml, argv[4] = signal directory
        if case == "stage":  # the s_stage flow: stage by stager.mode, dump, no scattering
            written, report = host.Job(sys.argv[3]).stage(comm=comm, backend=be.vtbl)
        else:
            written, report = host.Job(sys.argv[3]).run(sys.argv[4], comm=comm, backend=be.vtbl)
        gathered = [None] * world
        dist.all_gather_object(gathered, (rank, written, report))
        if rank == 0:
            with open(out_path, "wb") as f:
                pickle.dump(gathered, f)
        dist.barrier()
        dist.destroy_process_group()
        return
    NA, NF = 23, 12
    xyz = synth.trajectory(NF, NA, 20.0, 0.2, 3, offset=-10.0)
    b = synth.factors(NA)
    qv = host.create_from_scans([{"base": (1, 0, 0), "from": 0.3, "to": 1.5, "points": 5}])
    p = host.Params()
    if case.startswith("all"):
        p.set("scattering.average.orientation.type", "vectors").set("scattering.average.orientation.vectors.type", "file")
        p.set_vectors(synth.unit_vectors(7, 1))
    elif case.startswith("scan"):  # >= 8 subvectors: the runner batches the 5 |q| of the scan into one pass
        p.set("scattering.average.orientation.type", "vectors").set("scattering.average.orientation.vectors.type", "file")
        p.set_vectors(synth.unit_vectors(9, 1))
    elif case.startswith("self"):
        p.set("scattering.type", "self").set("scattering.average.orientation.type", "vectors")
        p.set("scattering.average.orientation.vectors.type", "file").set_vectors(synth.unit_vectors(3, 1))
    elif case.startswith("mp"):
        p.set("scattering.average.orientation.type", "multipole")
        p.set("scattering.average.orientation.multipole.moments.type", "resolution")
        p.set("scattering.average.orientation.multipole.moments.resolution", 3).set("scattering.dsp.type", "square")
    elif case.startswith("cyl"):
        p.set("scattering.average.orientation.type", "multipole").set("scattering.average.orientation.multipole.type", "cylinder")
        p.set("scattering.average.orientation.axis.x", 1).set("scattering.average.orientation.axis.z", 1)
        p.set("scattering.average.orientation.multipole.moments.type", "resolution")
        p.set("scattering.average.orientation.multipole.moments.resolution", 2)
    if case.endswith("_stream"):  # 12 frames x 12 B = 144 B per atom: two wave buffers of 5 atoms (config 5's streamed stager)
        p.set("limits.stage.memory.data", 10 * 144)
    if "_frames" in case:  # the reference's frame decomposition inside the partition (amplitude exchange)
        p.set("limits.decomposition.coherent", "frames")
    if case.endswith("_manual2"):  # partitions of two ranks: with three ranks one is left spare (scatter_device_factory.cpp:104-116)
        p.set("limits.decomposition.partitions.automatic", False).set("limits.decomposition.partitions.size", 2)
        p.set("limits.decomposition.utilization", 0.0)
    if case.endswith("_manual1"):  # partitions of one rank each: every rank owns a |q| subset, no all-reduce
        p.set("limits.decomposition.partitions.automatic", False).set("limits.decomposition.partitions.size", 1)
        p.set("limits.decomposition.utilization", 0.0)
    if case.endswith("_tight"):  # coordinate budget = 3/4 of the trajectory: fits a rank's HALF of the frames, not all of them
        p.set("limits.stage.memory.data", (NF * NA * 12 * 3) // 4)
    p.create()
    try:
        recs, has, tm = host.run_scatter(p, xyz, qv, b=b, comm=comm, backend=be.vtbl)
    except host.HostError as e:
        if not case.endswith("_tight"):
            raise
        recs, has, tm = str(e), False, {}
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, has, recs, sorted(tm)))
    if rank == 0:
        with open(out_path, "wb") as f:
            pickle.dump(gathered, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
