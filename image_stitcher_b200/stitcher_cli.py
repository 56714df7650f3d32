#!/usr/bin/env python3
"""``stitcher_cli`` -- the reference's synchronous CLI (stitcher_cli.py:87-116): same flags as ``stitcher_process_cli``,
but it builds a :class:`Stitcher` and calls ``run()`` in the calling process."""
from __future__ import annotations

import sys

from .stitcher_process_cli import create_params, parse_args


def main(argv=None) -> int:
    args = parse_args(argv)
    try:
        params = create_params(args)
        from .stitcher import Stitcher
        stitcher = Stitcher(params)
        print("Starting stitching with parameters:")
        for k in ("input_folder", "output_format", "apply_flatfield", "use_registration", "registration_channel",
                  "registration_z_level", "dynamic_registration", "scan_pattern"):
            print(f"{k.replace('_', ' ').capitalize()}: {getattr(params, k)}")
        stitcher.finished_saving.connect(lambda path, dtype: print(f"Stitching completed. Output saved to: {path}"))
        stitcher.run()
    except Exception as exc:
        print(f"Error: {exc}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
