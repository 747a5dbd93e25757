"""ORACLE helper: numpy restatement (float64 arithmetic, no FMA) of glibc's sincosf for |x| < 120, the same
algorithm csrc/orb_device.cuh runs on the device.  Used to check the restatement against libm."""
import numpy as np

_C = [float.fromhex(v) for v in ("0x1p0", "-0x1.ffffffd0c621cp-2", "0x1.55553e1068f19p-5",
                                 "-0x1.6c087e89a359dp-10", "0x1.99343027bf8c3p-16")]
_S = [float.fromhex(v) for v in ("-0x1.555545995a603p-3", "0x1.1107605230bc4p-7", "-0x1.994eb3774cf24p-13")]
_HPI_INV = float.fromhex("0x1.45F306DC9C883p+23")
_HPI = float.fromhex("0x1.921FB54442D18p0")


def _poly(x, x2, cs, n):
    sin_v = (x + (x * x2) * _S[0]) + ((x * x2) * x2) * (_S[1] + x2 * _S[2])
    x4 = x2 * x2
    cos_v = ((cs * _C[0] + x2 * (cs * _C[1])) + x4 * (cs * _C[2])) + (x4 * x2) * (cs * _C[3] + x2 * (cs * _C[4]))
    return np.where((n & 1) == 0, sin_v, cos_v).astype(np.float32)


def sincosf_restated(y: np.ndarray):
    y = np.asarray(y, np.float32)
    x = y.astype(np.float64)
    top = (y.view(np.uint32) >> 20) & 0x7FF
    pio4 = (np.float32(float.fromhex("0x1.921FB6p-1")).view(np.uint32) >> 20) & 0x7FF
    tiny = (np.float32(2.0 ** -12).view(np.uint32) >> 20) & 0x7FF
    out = []
    for which in (1, 0):   # cos, sin
        small = _poly(x, x * x, 1.0, np.full(x.shape, which))
        small = np.where(top < tiny, np.float32(1.0) if which else y, small)
        r = x * _HPI_INV
        n = (r.astype(np.int32) + 0x800000) >> 24
        xr = x - n * _HPI
        sgn = np.where(((n & 3) == 1) | ((n & 3) == 2), -1.0, 1.0)
        cs = np.where((n & 2) != 0, -1.0, 1.0)
        big = _poly(xr * sgn, xr * xr, cs, (n ^ 1) if which else n)
        out.append(np.where(top < pio4, small, big).astype(np.float32))
    return out[0], out[1]
