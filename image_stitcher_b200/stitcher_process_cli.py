#!/usr/bin/env python3
"""Command-line entry point -- same flags as the reference's ``stitcher_process_cli.py`` (:35-90), bound to
the CUDA-backed ``StitcherProcess`` of this package.

    python -m image_stitcher_b200.stitcher_process_cli -i DIR -r -ff --registration-channel "Fluorescence 488 nm Ex"

The worker runs as a separate ``multiprocessing.Process`` exactly like the reference's (fork on Linux): the
CUDA context is created lazily inside the worker's ``run()``, never in this parent.  It reports over the reference's three queues
(``('progress', (cur, total))``, ``('status', (msg, is_saving))``, ``('complete', (path, dtype))``,
``('error', msg)``) and honours the stop event on Ctrl-C (reference :113-185).
Extensions sit behind extra flags whose defaults reproduce the reference.
"""
from __future__ import annotations

import argparse
import dataclasses
import multiprocessing as mp
import sys
import time
from queue import Empty

from .stitcher_parameters import StitchingParameters


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Microscopy Image Stitching CLI (B200 hot path)")
    p.add_argument("--input-folder", "-i", required=True, help="Input folder containing images to stitch")
    p.add_argument("--output-format", "-f", choices=[".ome.zarr", ".ome.tiff"], default=".ome.zarr")
    p.add_argument("--apply-flatfield", "-ff", action="store_true", help="Apply flatfield correction")
    p.add_argument("--use-registration", "-r", action="store_true", help="Enable image registration")
    p.add_argument("--registration-channel", "-rc", help="Channel to use for registration (default: first channel)")
    p.add_argument("--registration-z-level", "-rz", type=int, default=0)
    p.add_argument("--dynamic-registration", action="store_true",
                   help="register every adjacent pair of the first region and use the median shifts")
    p.add_argument("--scan-pattern", "-s", choices=["Unidirectional", "S-Pattern"], default="Unidirectional")
    p.add_argument("--merge-timepoints", action="store_true")
    p.add_argument("--merge-hcs-regions", action="store_true")
    p.add_argument("--params-json", help="JSON file with stitching parameters (overrides the other arguments)")
    # extensions (defaults = reference behaviour)
    p.add_argument("--blend-mode", choices=["paste", "linear", "feather"], default="paste")
    p.add_argument("--placement", choices=["lattice", "global"], default="lattice",
                   help="lattice = the reference's single shift pair; global = all pairs of every region + least squares")
    p.add_argument("--upsample-factor", type=int, default=10)
    p.add_argument("--registration-precision", choices=["auto", "float32", "float64"], default="auto")
    p.add_argument("--visualize-registration", action="store_true",
                   help="write horizontal.png / vertical.png of the registered overlap strips (the reference always does)")
    p.add_argument("--device", type=int, default=0, help="CUDA device index")
    p.add_argument("--devices", default="", help="comma-separated CUDA devices: one worker process per device, regions "
                                                 "(wells) split round-robin, no inter-process exchange needed")
    p.add_argument("--split-regions", action="store_true",
                   help="with --devices: every worker fuses its (plane, chunk-row) bands of EVERY region into one shared "
                        "OME-Zarr (automatic when there are fewer regions than devices)")
    return p.parse_args(argv)


def create_params(args: argparse.Namespace) -> StitchingParameters:
    if args.params_json:
        return StitchingParameters.from_json(args.params_json)
    return StitchingParameters.from_dict({
        "input_folder": args.input_folder, "output_format": args.output_format,
        "apply_flatfield": args.apply_flatfield, "use_registration": args.use_registration,
        "registration_channel": args.registration_channel, "registration_z_level": args.registration_z_level,
        "scan_pattern": args.scan_pattern, "merge_timepoints": args.merge_timepoints,
        "merge_hcs_regions": args.merge_hcs_regions, "dynamic_registration": args.dynamic_registration,
        "blend_mode": args.blend_mode, "placement": args.placement, "upsample_factor": args.upsample_factor,
        "registration_precision": args.registration_precision, "device": args.device,
        "visualize_registration": args.visualize_registration, "split_regions": args.split_regions})


def monitor_process(proc, progress_queue, status_queue, complete_queue, stop_event, poll_s: float = 0.1) -> int:
    """Drain the three queues until the worker finishes; returns the process exit status (0 = ok)."""
    failed = False
    try:
        while True:
            busy = False
            for q in (progress_queue, status_queue, complete_queue):
                try:
                    kind, payload = q.get_nowait()
                except Empty:
                    continue
                busy = True
                if kind == "progress":
                    print(f"Progress: {payload[0]}/{payload[1]}")
                elif kind == "status":
                    print(f"Status: {payload[0]}")
                elif kind == "complete":
                    print(f"Stitching completed. Output saved to: {payload[0]}")
                elif kind == "error":
                    print(f"Error: {payload}", file=sys.stderr)
                    failed = True
            if not busy:
                if not proc.is_alive():
                    break
                time.sleep(poll_s)
    except KeyboardInterrupt:
        print("\nStopping stitching process...")
        stop_event.set()
        proc.join(timeout=3)
        if proc.is_alive():
            proc.terminate()
        return 130
    proc.join()
    return 1 if failed or proc.exitcode else 0


def main(argv=None) -> int:
    args = parse_args(argv)
    params = create_params(args)
    params.validate()
    from .stitcher_process import StitcherProcess
    devices = [int(d) for d in args.devices.split(",") if d.strip() != ""] or [params.device]
    _ = params.stitched_folder                      # freeze the output folder stamp once for all workers
    procs = []
    for rank, dev in enumerate(devices):
        p = dataclasses.replace(params, device=dev, rank=rank, world=len(devices))
        p._stamp = params._stamp
        queues = (mp.Queue(), mp.Queue(), mp.Queue())
        stop_event = mp.Event()
        proc = StitcherProcess(p, *queues, stop_event)
        proc.start()
        procs.append((proc, queues, stop_event))
    rc = 0
    for proc, queues, stop_event in procs:
        rc = max(rc, monitor_process(proc, *queues, stop_event))
    return rc


if __name__ == "__main__":
    sys.exit(main())
