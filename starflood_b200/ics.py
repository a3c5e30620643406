"""Synthetic inputs for bench.py and demos: a numpy restatement of the reference's initial-condition
generator (initcond.c:8-263, rng.h:13-40).  Input synthesis only - not on the measured path.  Positions and
masses agree with the reference bit for bit wherever numpy's libm calls round like glibc's (checked to 1 ulp
in tests/test_host_logic.py)."""
import numpy as np

PI_4_3 = 4.1887902047863909846168578443726705122628925325001410946332594564
PI_8_3 = 8.3775804095727819692337156887453410245257850650002821892665189128
TAU = 6.2831853071795864769252867665590057683943387987502116419498891846
INV_PCG32_MAX = 2.3283062977608182109179324470460414886474609375e-10


def pcg4d(s0, s1, s2, s3):
    """rng.h:13-40 on uint32 arrays (wrap-around arithmetic)."""
    v = [np.asarray(x, np.uint32).copy() for x in (s0, s1, s2, s3)]
    with np.errstate(over="ignore"):
        for k in range(4):
            v[k] = v[k] * np.uint32(0x0019660D) + np.uint32(0x3C6EF35F)
        for rnd in range(2):
            v[0] = v[0] + v[1] * v[3]
            v[1] = v[1] + v[2] * v[0]
            v[2] = v[2] + v[0] * v[1]
            v[3] = v[3] + v[1] * v[2]
            if rnd == 0:
                for k in range(4):
                    v[k] = v[k] ^ (v[k] >> np.uint32(16))
    return v


def _unit(u0, u1):
    phi = TAU * u0
    the = np.arccos(2.0 * u1 - 1.0)
    return np.cos(phi) * np.sin(the), np.sin(phi) * np.sin(the), np.cos(the)


def uniform_ball(N, outflow=4.0e-3, G=1.0):
    """(pos[3N], vel[3N], mas[N]) as initcond_generate() would fill them; outflow=0 drops the Hubble-like term
    (SURVEY 8d C3)."""
    i = np.arange(N, dtype=np.uint32)
    full = lambda c: np.full(N, c, np.uint32)
    with np.errstate(over="ignore"):
        s = pcg4d(np.uint32(0xD88D5EFD) + i, full(0x203F83D3), full(0x710BE493), full(0xA294E7F1))
    u = [INV_PCG32_MAX * x.astype(np.float64) for x in s]
    x, y, z = _unit(u[0], u[1])
    r = 1.0 * np.cbrt(u[2])
    pos = np.stack([x * r, y * r, z * r], axis=1).astype(np.float32)
    with np.errstate(over="ignore"):
        s = pcg4d(np.uint32(0x4AFDC7CF) + i, full(0xAE8D7CF2), full(0xE4827F00), full(0xE44CA389))
    u = [INV_PCG32_MAX * x.astype(np.float64) for x in s]
    n0 = np.sqrt(-2.0 * np.log(u[0])) * np.cos(TAU * u[1]) * 1.0e-4
    n1 = np.sqrt(-2.0 * np.log(u[0])) * np.sin(TAU * u[1]) * 1.0e-4
    n2 = np.sqrt(-2.0 * np.log(u[2])) * np.cos(TAU * u[3]) * 1.0e-4
    speed = np.sqrt((n0 * n0) + (n1 * n1) + (n2 * n2))
    s = pcg4d(*s)
    u = [INV_PCG32_MAX * x.astype(np.float64) for x in s]
    vx, vy, vz = _unit(u[0], u[1])
    p64 = pos.astype(np.float64)
    k = outflow * np.sqrt(PI_8_3 * G)
    vel = np.stack([vx * speed + k * p64[:, 0], vy * speed + k * p64[:, 1], vz * speed + k * p64[:, 2]], axis=1).astype(np.float32)
    mas = np.full(N, np.float32((PI_4_3 * 1.0) / N), np.float32)
    return pos.ravel(), vel.ravel(), mas


def galaxy_pair(N, G=1.0):
    """BASELINE.json C5: the author's disabled galaxy-pair sketch (initcond.c:265-329) on top of initcond_generate():
    85 % of the particles are squashed into a disc (y * 0.025) on circular-ish orbits, every particle joins one of two
    galaxies at x = +-2 approaching each other with +-0.005."""
    pos, vel, mas = uniform_ball(N, 4.0e-3, G)
    p = pos.reshape(-1, 3).astype(np.float64)
    v = vel.reshape(-1, 3).astype(np.float64)
    i = np.arange(N, dtype=np.uint32)
    full = lambda c: np.full(N, c, np.uint32)
    with np.errstate(over="ignore"):
        s = pcg4d(np.uint32(0x5D20CDF7) + i, full(0xE5C4E174), full(0xCB77B51E), full(0xA82C87A3))
    u = [INV_PCG32_MAX * x.astype(np.float64) for x in s]
    disc = u[0] < 0.850
    p[disc, 1] *= 0.025
    r2 = (p[:, 0] * p[:, 0]) + (p[:, 1] * p[:, 1]) + (p[:, 2] * p[:, 2])
    with np.errstate(divide="ignore"):
        inv_r1 = np.where(0.0 < r2, 1.0 / np.sqrt(r2), 0.0)
    k = 1.000e-2 * np.sqrt(PI_8_3 * G)
    v[disc, 0] = (k * inv_r1 * p[:, 2])[disc]
    v[disc, 1] = 0.0
    v[disc, 2] = (k * inv_r1 * -p[:, 0])[disc]
    first = u[1] < 0.500
    p[:, 0] += np.where(first, 2.000, -2.000)
    v[:, 0] += np.where(first, -0.005, 0.005)
    return p.astype(np.float32).ravel(), v.astype(np.float32).ravel(), mas
