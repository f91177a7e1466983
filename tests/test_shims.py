"""Known answers for oracle/shims ("a shim / real-library difference would go unnoticed").  The reference's sources are
compiled against minimal stand-ins for Eigen, KDL and tf2; this test drives the stand-ins directly (tests/cpp/test_shims.cpp) and
compares them with independent implementations: scipy.spatial.transform.Rotation for the matrix -> quaternion conversions (all four
branches of Shoemake's algorithm), the rotation-vector closed form of KDL::diff, the Hamilton product and the
angle formulas of tf2 LinearMath."""
import os
import subprocess

import numpy as np
from scipy.spatial.transform import Rotation

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "cpp", "test_shims")


def build():
    subprocess.run([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-ffp-contract=off", "-I" + os.path.join(ROOT, "oracle", "shims"), "-o", EXE, os.path.join(HERE, "cpp", "test_shims.cpp")], check=True)


def quats(n, rng):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    # rotations by ~pi about x, y and z: the three non-trace branches of matrix -> quaternion
    for k, axis in enumerate(np.eye(3)):
        for j in range(20):
            ang = np.pi - 1e-3 * j * (1 if j % 2 else -1)
            ax = axis + 0.02 * rng.normal(size=3)
            ax /= np.linalg.norm(ax)
            q[20 * k + j] = np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]])
    return q


def test_shims_against_scipy_and_closed_forms():
    build()
    rng = np.random.default_rng(11)
    n = 400
    q, r, p = quats(n, rng), quats(n, np.random.default_rng(12)), rng.normal(size=(n, 3))
    text = f"{n}\n" + "\n".join(" ".join(repr(float(x)) for x in np.concatenate([q[i], p[i], r[i]])) for i in range(n)) + "\n"
    out = subprocess.run([EXE], input=text, capture_output=True, text=True, check=True).stdout
    got = np.array([[float(x) for x in line.split()] for line in out.strip().split("\n")])
    assert got.shape == (n, 17)
    Rq, Rr = Rotation.from_quat(q), Rotation.from_quat(r)
    want = Rotation.from_matrix(Rq.as_matrix()).as_quat()

    def same_rotation(a, b):  # q and -q are the same rotation
        s = np.sign(np.sum(a * b, axis=1, keepdims=True))
        return np.abs(a * s - b).max()
    assert same_rotation(got[:, 0:4], want) < 1e-12  # Eigen::Quaterniond(Matrix3d)
    assert same_rotation(got[:, 4:8], want) < 1e-12  # KDL::Rotation::GetQuaternion
    traces = np.trace(Rq.as_matrix(), axis1=1, axis2=2)
    assert (traces > 0).sum() > 100 and (traces <= 0).sum() >= 60  # every branch was exercised
    # KDL::diff(R_a, R_b) = R_a * rotvec(R_a^T R_b)
    rel = (Rq.inv() * Rr).as_rotvec()
    assert np.abs(got[:, 8:11] - Rq.apply(rel)).max() < 1e-9
    # tf2::Quaternion operator* = Hamilton product
    assert same_rotation(got[:, 11:15], (Rq * Rr).as_quat()) < 1e-12 and np.abs(np.abs(np.sum(got[:, 11:15] * (Rq * Rr).as_quat(), axis=1)) - 1).max() < 1e-12
    # angleShortestPath = the rotation angle between the two orientations; Vector3::angle = acos of the normalised dot product
    assert np.abs(got[:, 15] - (Rq.inv() * Rr).magnitude()).max() < 1e-7
    v = np.array([1.0, 2.0, 3.0])
    assert np.abs(got[:, 16] - np.arccos(p @ v / np.linalg.norm(p, axis=1) / np.linalg.norm(v))).max() < 1e-12
