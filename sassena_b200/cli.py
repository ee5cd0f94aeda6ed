"""`python -m sassena_b200.cli --config scatter.xml` — the reference executable's command line for the GPU path
(reference src/main/sassena.cpp:200-260 and Params::options, src/control/parameters.cpp:795-836: --config and the overwrite
options --sample.structure.file / .format, --stager.target / .dump / .file / .format, --scattering.signal.file,
--limits.computation.threads).

One process per GPU: run under `python -m torch.distributed.run --nproc-per-node N -m sassena_b200.cli ...` for N GPUs.
The signal goes to scattering.signal.file (default signal.h5) as an HDF5 file in the reference's layout
(file_writer_service.cpp:44-171; an existing file is resumed like the reference does), with the per-rank rows kept as
.npy datasets under <file>.d/.  `--signal DIR` (no .h5 suffix) writes the .npy directory only.
`--stage-only` is the reference's second executable, s_stage (src/main/s_stage.cpp): stage and, with stager.dump, write the
post-processed trajectory."""
import argparse
import os
import sys


# Params::options (parameters.cpp:805-829)
OVERWRITES = [("sample.structure.file", "Structure file name"), ("sample.structure.format", "Structure file format"),
              ("stager.target", "Atom selection producing the signal (must be defined)"),
              ("stager.dump", "Do/Don't dump the postprocessed coordinates to a file"),
              ("stager.file", "Name of dump file"), ("stager.format", "Format of dump file"),
              ("scattering.signal.file", "name of the signal file"),
              ("limits.computation.threads", "Number of worker threads per process (no counterpart on the GPU path)")]


def main(argv=None):
    ap = argparse.ArgumentParser(prog="sassena_b200")
    ap.add_argument("--config", default="scatter.xml", help="xml configuration file (reference default: scatter.xml)")
    ap.add_argument("--signal", default=None, help="output file (.h5) or .npy directory (default: scattering.signal.file)")
    ap.add_argument("--device", type=int, default=None, help="CUDA device (default: LOCAL_RANK or 0)")
    ap.add_argument("--stage-only", action="store_true",
                    help="the reference's s_stage executable: stage the trajectory (stager.mode) and, with stager.dump, write "
                         "the post-processed coordinates to stager.file; no scattering calculation")
    ow = ap.add_argument_group("Overwrite options (applied after the configuration file has been read)")
    for key, text in OVERWRITES:
        ow.add_argument("--" + key, dest=key, default=None, metavar="ARG", help=text)
    args = ap.parse_args(argv)
    overwrites = {key: getattr(args, key) for key, _ in OVERWRITES if getattr(args, key) is not None}

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
    job = host.Job(args.config, overwrites)
    signal = args.signal
    if signal is None:
        signal = job.signal_file
    ctx = ScatterContext(dev)
    try:
        if args.stage_only:
            _, report = job.stage(comm=comm, ctx=ctx)
            signal = job.option("stager.file") if job.option("stager.dump") == "true" else "(no dump)"
        else:
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
