"""Host statement of the device simulator's noise function (eqvio_b200/csrc/simulator.cu: philox4x32_10 + normal_pair):
two independent N(0, 1) draws as a pure function of (seed, stream, event index, component).  Used by the tests to predict the
noise the device adds, and by anyone who wants the same Monte-Carlo instance on the host."""
import numpy as np

STREAM_IMU, STREAM_VISION = 1, 2
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c, k):
    """c: (..., 4) uint32 counters, k: (..., 2) uint32 keys -> (..., 4) uint32."""
    c = np.array(c, dtype=np.uint32, copy=True)
    k = np.array(np.broadcast_to(np.asarray(k, dtype=np.uint32), c.shape[:-1] + (2,)), copy=True)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[..., 0].astype(np.uint64)
            p1 = _M1 * c[..., 2].astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c[..., 1] ^ k[..., 0]
            n1 = p1.astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c[..., 3] ^ k[..., 1]
            n3 = p0.astype(np.uint32)
            c = np.stack([n0, n1, n2, n3], axis=-1)
            k = np.stack([k[..., 0] + _W0, k[..., 1] + _W1], axis=-1)
    return c


def normal_pair(seed, stream, event, comp):
    """(z0, z1) for scalar seed / stream and broadcastable integer arrays event, comp."""
    event = np.asarray(event, dtype=np.int64)
    comp = np.asarray(comp, dtype=np.int64)
    event, comp = np.broadcast_arrays(event, comp)
    ev = event.astype(np.uint64)
    c = np.stack([(ev & np.uint64(0xFFFFFFFF)).astype(np.uint32), (ev >> np.uint64(32)).astype(np.uint32), comp.astype(np.uint32),
                  np.full(event.shape, stream, dtype=np.uint32)], axis=-1)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    x = philox4x32_10(c, np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)).astype(np.uint64)
    a = ((x[..., 0] << np.uint64(32)) | x[..., 1]) >> np.uint64(11)
    b = ((x[..., 2] << np.uint64(32)) | x[..., 3]) >> np.uint64(11)
    u1 = (a.astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)
    u2 = b.astype(np.float64) * (1.0 / 9007199254740992.0)
    r = np.sqrt(-2.0 * np.log(u1))
    th = 6.283185307179586476925286766559 * u2
    return r * np.cos(th), r * np.sin(th)
