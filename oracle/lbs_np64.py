"""Oracle, second opinion: an INDEPENDENT float64 numpy formulation of SMPL / SMPL-X linear blend skinning.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **PARITY UNPINNED** exactly like oracle/lbs_ref.py: third-party
smplx==0.1.28 (reference requirements.txt:4) is absent, so neither file can be checked against it.  What this file
adds is that "unpinned" means "two restatements written differently agree to fp32 round-off":

  * lbs_ref.py follows the smplx source structure: 4x4 homogeneous matrices, the `A = G - pad(G @ Jrest)`
    correction, a dense W @ A matmul, fp32 torch.
  * this file uses the textbook form  v' = sum_j w_vj ( Rw_j (v_posed - Jrest_j) + Jposed_j )  with world rotations
    and posed joints propagated by recursion over the kinematic tree, no homogeneous matrices, fp64 numpy.

Both keep the two smplx-specific conventions the hot path depends on: Rodrigues with
angle = ||r + 1e-8|| and pose_feature = (R[1:] - I) in row-major order against posedirs [P, 3V].
Call sites in the reference: lib/body_model/body_model.py:75-88, lib/body_model/smpl.py:67-78.
"""
import numpy as np


def rodrigues(r):
    """[N,3] axis-angle -> [N,3,3]; angle = ||r + 1e-8|| as in smplx.lbs.batch_rodrigues."""
    r = np.asarray(r, np.float64)
    angle = np.sqrt(((r + 1e-8) ** 2).sum(1))
    k = r / angle[:, None]
    s, c = np.sin(angle), np.cos(angle)
    R = np.empty((r.shape[0], 3, 3))
    x, y, z = k[:, 0], k[:, 1], k[:, 2]
    C = 1.0 - c
    # I + s K + (1-c) K^2 written out element by element, keeping |k|^2 as smplx does (k is NOT exactly unit:
    # the 1e-8 shift enters the norm only), i.e. K^2 = k k^T - |k|^2 I
    n2 = x * x + y * y + z * z
    R[:, 0, 0] = 1 + C * (x * x - n2)
    R[:, 1, 1] = 1 + C * (y * y - n2)
    R[:, 2, 2] = 1 + C * (z * z - n2)
    R[:, 0, 1] = C * x * y - s * z
    R[:, 1, 0] = C * x * y + s * z
    R[:, 0, 2] = C * x * z + s * y
    R[:, 2, 0] = C * x * z - s * y
    R[:, 1, 2] = C * y * z - s * x
    R[:, 2, 1] = C * y * z + s * x
    return R


def body_forward(model, betas, full_pose, transl=None):
    """Same contract as oracle.lbs_ref.body_forward; returns float64 numpy (verts [B,V,3], joints [B,J+extra+lmk,3])."""
    f = lambda k: np.asarray(model[k], np.float64)              # noqa: E731
    vt, sdirs, pdirs, Jreg, W = f('v_template'), f('shapedirs'), f('posedirs'), f('J_regressor'), f('lbs_weights')
    parents = [int(p) for p in model['parents']]
    betas = np.asarray(betas, np.float64)
    pose = np.asarray(full_pose, np.float64)
    B, J, V = betas.shape[0], Jreg.shape[0], vt.shape[0]
    v_shaped = vt[None] + np.tensordot(betas, sdirs, axes=([1], [2]))           # [B,V,3]
    Jrest = np.einsum('jv,bvk->bjk', Jreg, v_shaped)
    R = rodrigues(pose.reshape(-1, 3)).reshape(B, J, 3, 3)
    feat = (R[:, 1:] - np.eye(3)).reshape(B, (J - 1) * 9)
    v_posed = v_shaped + (feat @ pdirs).reshape(B, V, 3)
    Rw = np.empty_like(R)
    Jp = np.empty_like(Jrest)
    Rw[:, 0], Jp[:, 0] = R[:, 0], Jrest[:, 0]
    for j in range(1, J):
        p = parents[j]
        Rw[:, j] = Rw[:, p] @ R[:, j]
        Jp[:, j] = Jp[:, p] + np.einsum('bik,bk->bi', Rw[:, p], Jrest[:, j] - Jrest[:, p])
    verts = np.zeros((B, V, 3))
    for j in range(J):
        w = W[:, j]
        nz = np.nonzero(w)[0]
        if nz.size == 0:
            continue
        local = v_posed[:, nz] - Jrest[:, j][:, None]
        verts[:, nz] += w[nz][None, :, None] * (np.einsum('bik,bvk->bvi', Rw[:, j], local) + Jp[:, j][:, None])
    parts = [Jp, verts[:, [int(i) for i in model['extra_vids']]]]
    if model.get('lmk_faces') is not None:
        faces = np.asarray(model['lmk_faces'], np.int64)
        bary = np.asarray(model['lmk_bary'], np.float64)
        parts.append(np.einsum('blfi,lf->bli', verts[:, faces.reshape(-1)].reshape(B, -1, 3, 3), bary))
    joints = np.concatenate(parts, axis=1)
    if transl is not None:
        t = np.asarray(transl, np.float64)[:, None]
        verts, joints = verts + t, joints + t
    return verts, joints
