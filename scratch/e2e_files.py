"""BASELINE.json configs[0] from FILES: synthetic 2x2 grid of 2048x2048 uint16 tiles in Squid layout -> registration +
stitch -> OME-Zarr, through the drop-in StitcherProcess (decode, GPU hot path, writer), timed per stage."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import synth
from image_stitcher_b200.stitcher_parameters import StitchingParameters
from image_stitcher_b200.stitcher_process import StitcherProcess
from image_stitcher_b200 import ome_zarr_writer as ozw

with tempfile.TemporaryDirectory() as tmp:
    root = os.path.join(tmp, "acq")
    st, tiles, truth = synth.make_region(rows=2, cols=2, tile_h=2048, tile_w=2048, seed=7, jitter=3, use_registration=True)
    synth.write_squid_layout(root, {"A1": tiles})
    s = StitcherProcess(StitchingParameters(input_folder=root, use_registration=True), None, None, None, None)
    t0 = time.perf_counter(); s.get_timepoints(); s.extract_acquisition_parameters(); s.get_pixel_size(); s.parse_acquisition_metadata()
    t1 = time.perf_counter(); _ = s.ctx; t2 = time.perf_counter()
    s.calculate_shifts(s.timepoints[0], s.regions[0]); t3 = time.perf_counter()
    out = s.stitch_region(0, "A1"); t4 = time.perf_counter()
    s.calculate_shifts(s.timepoints[0], s.regions[0]); t5 = time.perf_counter()
    out = s.stitch_region(0, "A1"); t6 = time.perf_counter()
    os.makedirs(os.path.join(s.output_folder, "0_stitched"), exist_ok=True)
    path = s.save_region_ome_zarr(0, "A1", out); t7 = time.perf_counter()
    ok = np.array_equal(ozw.read_ome_zarr_level(path, 0), out)
    print(f"shifts h={s.h_shift} v={s.v_shift} truth h={truth['h_shift']} v={truth['v_shift']} canvas={out.shape} zarr_roundtrip={ok}")
    print(f"parse {t1-t0:.3f}s | CUDA context {t2-t1:.3f}s | calculate_shifts first {t3-t2:.3f}s warm {t5-t4:.3f}s (4 TIFF decodes + 2 pairs) | "
          f"stitch_region first {t4-t3:.3f}s warm {t6-t5:.3f}s ({out.size/1e6:.1f} Mpx, 4 TIFF decodes) | OME-Zarr write {t7-t6:.3f}s")
    s.cleanup()
