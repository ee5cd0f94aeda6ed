"""`python -m sassena_b200.cli --config scatter.xml` — the reference executable's command line for the GPU path
(reference src/main/sassena.cpp:200-260: --config, and the scattering.signal.file override).

One process per GPU: run under `python -m torch.distributed.run --nproc-per-node N -m sassena_b200.cli ...` for N GPUs.
The signal is written as a directory of .npy datasets named like the reference's signal.h5 datasets."""
import argparse
import os
import sys


def main(argv=None):
    ap = argparse.ArgumentParser(prog="sassena_b200")
    ap.add_argument("--config", default="scatter.xml", help="xml configuration file (reference default: scatter.xml)")
    ap.add_argument("--signal", default=None, help="output directory (default: scattering.signal.file with .npy.d suffix)")
    ap.add_argument("--device", type=int, default=None, help="CUDA device (default: LOCAL_RANK or 0)")
    args = ap.parse_args(argv)

    from . import host
    from .api import ScatterContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = args.device if args.device is not None else local
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        comm = host.TorchDistCommunicator()
    job = host.Job(args.config)
    signal = args.signal
    if signal is None:
        base = os.path.splitext(os.path.basename(args.config))[0]
        signal = os.path.join(os.path.dirname(os.path.abspath(args.config)), base + ".signal.d")
    ctx = ScatterContext(dev)
    try:
        written, report = job.run(signal, comm=comm, ctx=ctx)
    finally:
        ctx.close()
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"sassena_b200: {report} -> {signal}", file=sys.stderr)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
