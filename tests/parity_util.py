"""Shared helpers for the parity tests: oracle state <-> the C ABI's flat qpos/qvel/aux arrays (include/plen_b200.h)."""
import numpy as np


def abi_from_oracle(st):
    """dict from PlenOracle.get_state() -> (qpos[N,25], qvel[N,24], aux[N,29]) float32"""
    n = st["qpos"].shape[0]
    aux = np.zeros((n, 29), dtype=np.float32)
    aux[:, 0:8] = st["lam_n"]
    aux[:, 8] = (st["in_manifold"].astype(np.int64) << np.arange(8)).sum(1)
    aux[:, 9:13] = st["book_i"][:, 0:4]
    aux[:, 13:19] = st["book_f"][:, 0:6]
    aux[:, 19:28] = st["book_f"][:, 6:15]
    aux[:, 28] = st["book_f"][:, 15]
    return st["qpos"].astype(np.float32), st["qvel"].astype(np.float32), aux


def oracle_from_abi(qpos, qvel, aux):
    """inverse of abi_from_oracle (float64 dict for PlenOracle.set_state); `dead` and targets are left out"""
    qpos, qvel, aux = (np.asarray(a, dtype=np.float64) for a in (qpos, qvel, aux))
    n = qpos.shape[0]
    st = dict(qpos=qpos.copy(), qvel=qvel.copy(), lam_n=aux[:, 0:8].copy())
    st["in_manifold"] = ((aux[:, 8:9].astype(np.int64) >> np.arange(8)) & 1).astype(np.int32)
    bi = np.zeros((n, 5), dtype=np.int32)
    bi[:, 0:4] = aux[:, 9:13]
    st["book_i"] = bi
    bf = np.zeros((n, 16))
    bf[:, 0:6] = aux[:, 13:19]
    bf[:, 6:15] = aux[:, 19:28]
    bf[:, 15] = aux[:, 28]
    st["book_f"] = bf
    return st


def random_flight_state(rng, n):
    """Airborne robots (z = 0.5 m, no contact) with random pose and velocities, as oracle-state dict pieces."""
    qpos = np.zeros((n, 25))
    qvel = np.zeros((n, 24))
    qpos[:, 0:2] = rng.uniform(-0.3, 0.3, (n, 2))
    qpos[:, 2] = 0.5
    q = rng.normal(size=(n, 4))
    qpos[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    qpos[:, 7:] = rng.uniform(-0.6, 0.6, (n, 18))
    qvel[:, 0:6] = rng.normal(size=(n, 6)) * 0.3
    qvel[:, 6:] = rng.normal(size=(n, 18)) * 2.0
    return qpos, qvel
