"""
Rotation-representation converters used at the parameter pack/unpack boundary.

The residual itself uses Euler angles with R = Rz(yaw) @ Ry(pitch) @ Rx(roll)
(reference: bundle_adjust/ba_core.py:36-56, bundle_adjust/ba_rotate.py:67-94).
Only the two converters the BA path calls are provided; quaternion / axis-angle
helpers of the reference are unused by BA and out of scope.
"""
import numpy as np


def euler_angles_from_R(R):
    """3x3 rotation -> (roll, pitch, yaw); same branch rule as ba_rotate.py:67-82."""
    sy = np.sqrt(R[0, 0] * R[0, 0] + R[1, 0] * R[1, 0])
    pitch = np.arctan2(-R[2, 0], sy)
    if sy < 1e-6:  # gimbal lock: yaw is not observable, fold it into roll
        return np.arctan2(-R[1, 2], R[1, 1]), pitch, 0
    return np.arctan2(R[2, 1], R[2, 2]), pitch, np.arctan2(R[1, 0], R[0, 0])


def euler_angles_to_R(roll, pitch, yaw):
    """(roll, pitch, yaw) -> Rz @ Ry @ Rx (ba_rotate.py:85-94)."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cw, sw = np.cos(yaw), np.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    Rz = np.array([[cw, -sw, 0], [sw, cw, 0], [0, 0, 1]], dtype=np.float64)
    return Rz @ Ry @ Rx
