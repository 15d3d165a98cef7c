"""An INDEPENDENTLY DERIVED model of the Vine5LinkMovingBase mechanism, for anchoring the oracle's dynamics.

The oracle (oracle/vine_oracle_dyn.inc) and the CUDA kernel share hand-derived closed-form equations of motion
(absolute link angles, beta/alpha coefficient recursion, semi-implicit Euler).  Nothing of that is used here:
this file states only the URDF's geometry and inertias (assets/urdf/Vine5LinkMovingBase.urdf:56-87,
264-326) as Cartesian forward kinematics in the reference's RELATIVE joint coordinates, builds the Lagrangian
L = T - V from it, and lets torch autograd produce the equations of motion:

    M(q) qdd = Q + dL/dq - (d^2L/dqd dq) qd,      M = d^2L/dqd^2

integrated in continuous time with scipy's DOP853.  Test infrastructure only.
"""
import numpy as np
import torch
from scipy.integrate import solve_ivp
from torch.func import hessian, jacrev

# URDF
LINK_LEN, LINK_COM = 0.0885, 0.04425               # joint-to-joint distance, COM offset along the link (URDF:294-321, 81-87)
BASE_ANGLE = 3.1415                                # rpy of the first revolute joint origin (URDF:289), NOT pi
PIVOT_Z = 1.0 - 0.025 - 0.01                       # slider z=1.0, cart joint -0.025, link_0 joint -0.01 (URDF:270-293)
CART_MASS = 0.4
LINK_MASS = (0.005, 0.005, 0.005, 0.005, 0.1)
LINK_INERTIA = (6.89246e-6, 6.89246e-6, 6.89246e-6, 6.89246e-6, 1.01559e-4)   # I_xx about each COM
GRAVITY = 9.81
# torque law V5:1045-1048
TL_K = np.array([0.8385, 1.5400, 1.5109, 1.2887, 0.4347])
TL_C = np.array([0.0178, 0.0304, 0.0528, 0.0367, 0.0223])
TL_b = np.array([0.0007, 0.0062, 0.0402, 0.0160, 0.0133])
TL_B = np.array([0.0247, 0.0616, 0.0779, 0.0498, 0.0268])


def joint_points(q):
    """(y,z) of the 6 chain points p_0 (first revolute joint) .. p_5 (tip) and the absolute link angles."""
    phi = BASE_ANGLE + torch.cumsum(q[1:], 0)
    dy, dz = -torch.sin(phi), torch.cos(phi)        # a link at angle phi about +x points along R_x(phi) e_z
    py, pz = [q[0]], [torch.as_tensor(PIVOT_Z, dtype=q.dtype)]
    for k in range(5):
        py.append(py[-1] + LINK_LEN * dy[k]); pz.append(pz[-1] + LINK_LEN * dz[k])
    return torch.stack(py), torch.stack(pz), phi


def lagrangian(q, qd):
    py, pz, phi = joint_points(q)
    dy, dz = -torch.sin(phi), torch.cos(phi)
    comy, comz = py[:5] + LINK_COM * dy, pz[:5] + LINK_COM * dz
    # velocities by the chain rule through autograd-friendly directional derivatives
    Jy = jacrev(lambda qq: _com(qq)[0])(q)          # [5,6]
    Jz = jacrev(lambda qq: _com(qq)[1])(q)
    vy, vz = Jy @ qd, Jz @ qd
    w = torch.cumsum(qd[1:], 0)
    m = torch.tensor(LINK_MASS, dtype=q.dtype); inertia = torch.tensor(LINK_INERTIA, dtype=q.dtype)
    T = 0.5 * CART_MASS * qd[0] ** 2 + 0.5 * (m * (vy ** 2 + vz ** 2)).sum() + 0.5 * (inertia * w ** 2).sum()
    V = GRAVITY * (m * comz).sum()
    del comy
    return T - V


def _com(q):
    py, pz, phi = joint_points(q)
    return py[:5] - LINK_COM * torch.sin(phi), pz[:5] + LINK_COM * torch.cos(phi)


def kinetic(q, qd):
    return lagrangian(q, qd) + potential(q)


def potential(q):
    m = torch.tensor(LINK_MASS, dtype=q.dtype)
    return GRAVITY * (m * _com(q)[1]).sum()


def mass_matrix(q):
    q = torch.as_tensor(q, dtype=torch.float64)
    return hessian(lambda qd: lagrangian(q, qd))(torch.zeros(6, dtype=torch.float64)).numpy()


def acceleration(q, qd, Q):
    """qdd from the Euler-Lagrange equations with generalized forces Q (numpy in, numpy out)."""
    q = torch.as_tensor(q, dtype=torch.float64); qd = torch.as_tensor(qd, dtype=torch.float64)
    M = hessian(lagrangian, argnums=1)(q, qd)
    dLdq = jacrev(lagrangian, argnums=0)(q, qd)
    mixed = jacrev(jacrev(lagrangian, argnums=1), argnums=0)(q, qd)   # d/dq (dL/dqd): [6,6]
    rhs = torch.as_tensor(Q, dtype=torch.float64) + dLdq - mixed @ qd
    return torch.linalg.solve(M, rhs).numpy()


def integrate(q0, qd0, duration, force_fn, rtol=1e-10, atol=1e-12):
    """Continuous-time solution; force_fn(q, qd) -> generalized forces [6] (numpy)."""
    def rhs(t, s):
        q, qd = s[:6], s[6:]
        return np.concatenate([qd, acceleration(q, qd, force_fn(q, qd))])
    sol = solve_ivp(rhs, (0.0, duration), np.concatenate([q0, qd0]), method="DOP853", rtol=rtol, atol=atol)
    assert sol.success
    return sol.y[:6, -1], sol.y[6:, -1]


def energy(q, qd, spring_k=None):
    q = torch.as_tensor(q, dtype=torch.float64); qd = torch.as_tensor(qd, dtype=torch.float64)
    e = float(kinetic(q, qd) + potential(q))
    if spring_k is not None:
        e += float(0.5 * (np.asarray(spring_k) * q[1:].numpy() ** 2).sum())
    return e


def equilibrium(u=0.0, iters=60):
    """Static equilibrium of the revolutes under gravity and the torque law tau = -(K q + b + B u) (cart anywhere)."""
    q = torch.zeros(6, dtype=torch.float64)
    K, b, B = (torch.tensor(x, dtype=torch.float64) for x in (TL_K, TL_b, TL_B))

    def residual(qr):
        qq = torch.cat([torch.zeros(1, dtype=torch.float64), qr])
        g = jacrev(potential)(qq)[1:]
        return g + K * qr + b + B * u

    qr = q[1:].clone()
    for _ in range(iters):
        r = residual(qr)
        J = jacrev(residual)(qr)
        qr = qr - torch.linalg.solve(J, r)
    assert float(residual(qr).abs().max()) < 1e-12
    return np.concatenate([[0.0], qr.numpy()])


def modal_frequencies(q_eq):
    """Undamped small-oscillation frequencies (rad/s) about q_eq with the torque-law stiffness, cart free."""
    import scipy.linalg
    M = mass_matrix(q_eq)
    Kg = hessian(potential)(torch.as_tensor(q_eq, dtype=torch.float64)).numpy()
    Kt = Kg + np.diag(np.concatenate([[0.0], TL_K]))
    w2, vec = scipy.linalg.eigh(Kt, M)
    return np.sqrt(np.clip(w2, 0, None)), vec, M
