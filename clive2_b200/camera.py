"""Sensor-plus-pinhole camera and the log-average tone map.

Behavioural mirror of the reference's src/camera.py:7-82.  The camera is a physical sensor rectangle centred on
`center` with the pinhole (`focal_point`) `focal_dist` in front of it along `direction`; the direction is stored
exactly as given (the presets pass un-normalised vectors, scene.py:156) because the kernels use it as the sensor
normal (trace.metal:610,1049).
"""
import numpy as np

from .constants import H_FOV, UNIT_X, UNIT_Y, UNIT_Z
from .struct_types import Camera as CameraRecord

_AXIS_EPS = 0.0001


class Camera:
    def __init__(self, center=np.zeros(3), direction=np.array([1, 0, 0]), phys_width=1.0, phys_height=1.0,
                 pixel_width=1280, pixel_height=720):
        self.center = center
        self.direction = direction
        self.phys_width = phys_width
        self.phys_height = phys_height
        self.pixel_width = pixel_width
        self.pixel_height = pixel_height
        self.aspect_ratio = phys_width / phys_height
        self.h_fov = H_FOV
        # camera.py:23
        self.v_fov = 2.0 * np.arctan(np.tan(H_FOV / 2.0) / self.aspect_ratio)

        dx, dy = self.dx, self.dy
        self.dx_dp = dx * self.phys_width / self.pixel_width
        self.dy_dp = dy * self.phys_height / self.pixel_height
        self.pixel_phys_size = np.linalg.norm(self.dx_dp) * np.linalg.norm(self.dy_dp)
        # lower corner of the sensor rectangle (camera.py:31); load.camera_geometry builds the two sensor triangles from it
        self.origin = center - dx * phys_width / 2 - dy * phys_height / 2

    # -- derived quantities (camera.py:33-55) -------------------------------------------------------------------
    @property
    def focal_dist(self):
        return self.phys_width / (2 * np.tan(self.h_fov / 2))

    @property
    def focal_point(self):
        return self.center + self.focal_dist * self.direction

    @property
    def dx(self):
        d = self.direction
        if abs(d[0]) < _AXIS_EPS:
            return UNIT_X if d[2] > 0 else -1 * UNIT_X
        horizontal = np.cross(d * (UNIT_X + UNIT_Z), -1 * UNIT_Y)
        return horizontal / np.linalg.norm(horizontal)

    @property
    def dy(self):
        d = self.direction
        if abs(d[1]) < _AXIS_EPS:
            return UNIT_Y
        up = np.cross(d, self.dx)
        return up / np.linalg.norm(up)

    def to_struct(self):
        """One-element `struct_types.Camera` array, the kernels' `camera[0]` (camera.py:57-70)."""
        rec = np.zeros(1, dtype=CameraRecord)
        r = rec[0]
        for name, vec in (("center", self.center), ("focal_point", self.focal_point), ("direction", self.direction),
                          ("dx", self.dx), ("dy", self.dy)):
            r[name][:3] = vec
        for name in ("pixel_width", "pixel_height", "phys_width", "phys_height", "h_fov", "v_fov"):
            r[name] = getattr(self, name)
        return rec


_LUMA_BGR = np.array([0.0722, 0.7152, 0.2126])


def tone_map(image, exposure=2.0, white_point=1.0, verbose=False):
    """Log-average-luminance tone map to uint8 BGR (camera.py:73-82)."""
    if verbose:
        print(f"IN min: {np.min(image)}, mean: {np.mean(image)}, max: {np.max(image)}")
    luminance = (image * _LUMA_BGR).sum(axis=2)
    log_avg = np.exp(np.log(0.1 + luminance).sum() / (image.shape[0] * image.shape[1]))
    scaled = image * exposure / log_avg
    return (255 * scaled / (scaled + white_point ** 2)).astype(np.uint8)


def basic_tone_map(image):
    return (255 * np.sqrt(image) / image).astype(np.uint8)
