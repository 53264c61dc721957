"""Record layouts shared by the host code, the C ABI and the CUDA kernels.

These are the six numpy structured dtypes that make up the reference's API surface
(reference: src/struct_types.py:4-85, mirrored by the MSL structs at src/trace.metal:7-85).
The byte layout is the contract: every float3 occupies a 16-byte slot (4 floats), scalar
members follow, and each record is padded to a multiple of 16 bytes.  `include/clive_b200.h`
declares the same layouts for C callers and `tests/test_layouts.py` pins the offsets against
the golden values obtained by importing the reference module (tests/golden/layouts.json).

The dtypes are generated from a small field table instead of being spelled out one by one so
that the C header generator in `tools/gen_header_check.py` and the tests can iterate over it.
"""
import numpy as np

_F3 = ("f3",)  # float3 stored in a 4-float slot
_F = ("f",)
_I = ("i",)


def _pad(n):
    return ("ipad", n)


# name -> ordered (field, kind) table
LAYOUTS = {
    "Ray": [
        ("origin", _F3), ("direction", _F3), ("inv_direction", _F3), ("color", _F3), ("normal", _F3),
        ("material", _I), ("triangle", _I),
        ("c_importance", _F), ("l_importance", _F), ("tot_importance", _F),
        ("hit_light", _I), ("from_camera", _I), ("hit_camera", _I), ("pixel_idx", _I),
        ("pad", _pad(3)),
    ],
    "Box": [("min", _F3), ("max", _F3), ("left", _I), ("right", _I), ("pad", _pad(2))],
    "Triangle": [
        ("v0", _F3), ("v1", _F3), ("v2", _F3), ("n0", _F3), ("n1", _F3), ("n2", _F3), ("normal", _F3),
        ("material", _I), ("is_light", _I), ("is_camera", _I), ("pad", _I),
    ],
    "Material": [("color", _F3), ("emission", _F3), ("type", _I), ("alpha", _F), ("ior", _F), ("pad", _I)],
    "Camera": [
        ("center", _F3), ("focal_point", _F3), ("direction", _F3), ("dx", _F3), ("dy", _F3),
        ("pixel_width", _I), ("pixel_height", _I),
        ("phys_width", _F), ("phys_height", _F), ("h_fov", _F), ("v_fov", _F),
        ("pad", _pad(2)),
    ],
}


def _np_field(name, kind):
    tag = kind[0]
    if tag == "f3":
        return (name, np.float32, 4)
    if tag == "f":
        return (name, np.float32)
    if tag == "i":
        return (name, np.int32)
    if tag == "ipad":
        return (name, np.int32, kind[1])
    raise ValueError(tag)


def _build(name):
    return np.dtype([_np_field(f, k) for f, k in LAYOUTS[name]])


Ray = _build("Ray")
Box = _build("Box")
Triangle = _build("Triangle")
Material = _build("Material")
Camera = _build("Camera")

#: number of vertex slots in a Path record (reference: renderer.py:8 MAX_PATH_LENGTH, trace.metal:32)
PATH_SLOTS = 8
#: bounce limit actually used by the random walk (reference: trace.metal:407 `i < 6`)
MAX_BOUNCES = 6

Path = np.dtype([("rays", Ray, PATH_SLOTS), ("length", np.int32), ("from_camera", np.int32), ("pad", np.int32, 2)])

#: internal per-sample record of connect_paths/adaptive_finalize_samples (trace.metal:25-29); the
#: reference allocates it with a 128-byte stride (renderer.py:71) although MSL's sizeof is 80.
WeightAggregator = np.dtype(
    {
        "names": ["weights", "total_contribution", "contrib_weight_sum"],
        "formats": [(np.float32, (3, 3)), (np.float32, 4), np.float32],
        "offsets": [0, 48, 64],
        "itemsize": 80,
    }
)

EXPECTED_SIZES = {"Ray": 128, "Path": 1040, "Box": 48, "Triangle": 128, "Material": 48, "Camera": 112}

for _n, _sz in EXPECTED_SIZES.items():
    assert globals()[_n].itemsize == _sz, (_n, globals()[_n].itemsize)
