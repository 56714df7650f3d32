"""Configuration of a stitching run -- same flag surface as the reference's
``stitcher_parameters.StitchingParameters`` (stitcher_parameters.py:8-107) so CLIs, JSON files
and GUIs written against the reference keep working, plus optional knobs for the extensions
(all defaulting to the reference's behaviour).
"""
from __future__ import annotations

import dataclasses
import datetime
import json
import os
from typing import Any, Dict

_FORMATS = (".ome.zarr", ".ome.tiff")
_PATTERNS = ("Unidirectional", "S-Pattern")
_BLENDS = ("paste", "linear", "feather")
_PRECISIONS = ("auto", "float32", "float64")


@dataclasses.dataclass
class StitchingParameters:
    # --- the reference's fields (names, defaults and meaning unchanged) ---
    input_folder: str
    output_format: str = ".ome.zarr"
    apply_flatfield: bool = False
    use_registration: bool = False
    registration_channel: str = ""        # empty -> first channel found
    registration_z_level: int = 0
    dynamic_registration: bool = False    # declared-but-unused in the reference; here: all-pairs registration
    scan_pattern: str = "Unidirectional"
    merge_timepoints: bool = False
    merge_hcs_regions: bool = False
    # --- extensions (defaults reproduce the reference) ---
    blend_mode: str = "paste"             # 'paste' = reference crop-to-seam + overwrite
    upsample_factor: int = 10             # the reference hard-codes 10 (stitcher_process.py:684)
    registration_precision: str = "auto"
    placement: str = "lattice"            # 'lattice' = the reference's single (h_shift, v_shift) model; 'global' = every
                                          # adjacent pair of every region registered, least-squares tile positions
    visualize_registration: bool = False  # write <out>/horizontal.png / vertical.png like the reference's visualize_image
                                          # side effect (stitcher_process.py:681, 704, 857-881); off: it costs a host round trip
    device: int = 0
    rank: int = 0                         # multi-GPU: this worker stitches regions rank, rank + world, ...
    world: int = 1
    split_regions: bool = False           # multi-GPU: every worker fuses its (plane, chunk-row) bands of EVERY region
                                          # (automatic when there are fewer regions than workers)

    def __post_init__(self):
        self.input_folder = os.path.abspath(self.input_folder)
        self._stamp = None

    def validate(self) -> None:
        problems = []
        if not os.path.exists(self.input_folder):
            problems.append(f"Input folder does not exist: {self.input_folder}")
        if self.output_format not in _FORMATS:
            problems.append("Output format must be either .ome.zarr or .ome.tiff")
        if self.scan_pattern not in _PATTERNS:
            problems.append("Scan pattern must be either 'Unidirectional' or 'S-Pattern'")
        if self.use_registration and self.registration_z_level < 0:
            problems.append("Registration Z-level must be non-negative")
        if self.placement not in ("lattice", "global"):
            problems.append("placement must be 'lattice' or 'global'")
        if self.blend_mode not in _BLENDS:
            problems.append(f"blend_mode must be one of {_BLENDS}")
        if self.registration_precision not in _PRECISIONS:
            problems.append(f"registration_precision must be one of {_PRECISIONS}")
        if not 1 <= int(self.upsample_factor) <= 100:
            problems.append("upsample_factor must be in [1, 100]")
        if problems:
            raise ValueError("; ".join(problems))

    @property
    def stitched_folder(self) -> str:
        """``<input>_stitched_<timestamp>``.  The reference re-evaluates ``datetime.now()`` on every access
        (stitcher_parameters.py:62-64, a defect noted in SURVEY.md 2.3); the stamp is frozen on first use here."""
        if getattr(self, "_stamp", None) is None:
            self._stamp = datetime.datetime.now().strftime("%Y-%m-%d_%H-%M-%S.%f")
        return f"{self.input_folder}_stitched_{self._stamp}"

    # ---- (de)serialisation, unknown keys ignored like the reference does
    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "StitchingParameters":
        names = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in data.items() if k in names and v is not None})

    @classmethod
    def from_json(cls, json_path: str) -> "StitchingParameters":
        with open(json_path) as fh:
            return cls.from_dict(json.load(fh))

    def to_dict(self) -> Dict[str, Any]:
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}

    def to_json(self, json_path: str) -> None:
        with open(json_path, "w") as fh:
            json.dump(self.to_dict(), fh, indent=2)
