"""The kernels' __host__ __device__ arithmetic (csrc/foho_math.cuh), compiled for the host
and compared with the oracle -- including the bit-exact inside/outside rule."""
import ctypes as C

import numpy as np
import pytest
import torch

from followmyhold_b200.synthetic import cap_boundary_loops, icosphere, random_similarity, standin_hand_mesh
from oracle import guidance_oracle as O


@pytest.fixture(scope="module")
def host():
    import __graft_entry__ as g
    lib = C.CDLL(g.build_host_check())
    lib.host_closest_point.restype = C.c_float
    return lib


def fp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_quaternion_matrix_and_backward(host):
    rng = np.random.default_rng(0)
    for _ in range(10):
        q = rng.normal(size=4).astype(np.float32) * rng.uniform(0.3, 3)
        GR = rng.normal(size=9).astype(np.float32)
        R = np.zeros(9, np.float32); gq = np.zeros(4, np.float32)
        host.host_quat_to_mat(fp(q), fp(R)); host.host_quat_backward(fp(q), fp(GR), fp(gq))
        qt = torch.tensor(q, dtype=torch.float64, requires_grad=True)
        Rt = O.quaternion_to_matrix(qt).reshape(-1)
        (Rt * torch.tensor(GR, dtype=torch.float64)).sum().backward()
        assert np.abs(R - Rt.detach().numpy()).max() < 1e-6
        assert np.abs(gq - qt.grad.numpy()).max() < 2e-5 * max(1.0, np.abs(qt.grad.numpy()).max())


def test_closest_point_on_triangle(host):
    rng = np.random.default_rng(1)
    for _ in range(3000):
        tri = rng.normal(size=(3, 3)).astype(np.float32)
        p = (rng.normal(size=3) * 1.5).astype(np.float32)
        w = np.zeros(3, np.float32)
        d2 = host.host_closest_point(fp(p), fp(tri[0]), fp(tri[1]), fp(tri[2]), fp(w))
        od2, _, ob = O.closest_point_barycentric(torch.tensor(p[None], dtype=torch.float64), torch.tensor(tri[None], dtype=torch.float64))
        assert abs(d2 - float(od2)) < 1e-5 * max(1.0, float(od2))
        assert abs(w.sum() - 1) < 1e-5 and (w >= -1e-6).all()
        q = (w[:, None] * tri).sum(0)
        assert abs(((p - q) ** 2).sum() - d2) < 1e-5 * max(1.0, d2)


@pytest.mark.parametrize("capped", [False, True])
def test_raster_rule_is_bit_exact_with_the_oracle(host, capped):
    v, f = standin_hand_mesh()
    if capped:
        f = cap_boundary_loops(f)
    D = 40
    for seed in range(4):
        T = random_similarity(seed, (50, 90), 0.0)
        hg = (v.astype(np.float64) @ T[:3, :3].T + np.array([20, 19.5, 20.25])).astype(np.float32)
        out = np.zeros(D ** 3, np.uint8)
        host.host_raster_parity(fp(hg), fp(np.ascontiguousarray(f)), len(f), D, fp(out))
        ref = O.raster_parity_inside(hg, f, D)
        assert (out.reshape(D, D, D).astype(bool) == ref).all()
        assert ref.sum() > 100
    # integer-coordinate vertices: exact ties on edges and vertices, still identical
    sv, sf = icosphere(1, 8.0)
    sv = np.round(sv + 16).astype(np.float32)
    out = np.zeros(33 ** 3, np.uint8)
    host.host_raster_parity(fp(sv), fp(np.ascontiguousarray(sf)), len(sf), 33, fp(out))
    assert (out.reshape(33, 33, 33).astype(bool) == O.raster_parity_inside(sv, sf, 33)).all()


def test_kabsch_rotation_matches_numpy_svd(host):
    rng = np.random.default_rng(2)
    for k in range(20):
        H = rng.normal(size=(3, 3))
        if k % 4 == 0:
            H[:, 2] *= 1e-3          # nearly planar configuration
        if k % 5 == 0:
            H = H @ np.diag([1, 1, -1.0])   # would need a reflection: must still return det +1
        R = np.zeros(9)
        host.host_kabsch(fp(np.ascontiguousarray(H.reshape(-1))), fp(R))
        u, s, vh = np.linalg.svd(H)
        Rr = u @ np.diag([1, 1, np.linalg.det(u @ vh)]) @ vh
        assert np.abs(R.reshape(3, 3) - Rr).max() < 1e-9
        assert abs(np.linalg.det(R.reshape(3, 3)) - 1) < 1e-12
