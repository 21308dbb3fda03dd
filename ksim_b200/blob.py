"""Flattens a compiled ``Model`` (+ task program) into the byte blob ``b200sim_model_create`` uploads.

Layout: ``include/b200sim_model.h``.  This is the device-side counterpart of ``mjx.put_model``
(ksim/utils/mujoco.py:388-391): the arrays a rollout needs, in fp32 / int32, plus a few derived index
tables (tree levels, subtree ranges, dof chains) that let one warp walk the kinematic tree without recursion.
"""

from __future__ import annotations

import numpy as np

from ksim_b200 import codes as C
from ksim_b200.mjcf import GEOM, JNT, Model

__all__ = ["build_model_arrays", "pack_blob", "contacts_per_pair"]


def contacts_per_pair(model: Model) -> list[int]:
    out = []
    for g1, g2 in zip(model.pair_geom1, model.pair_geom2):
        t1, t2 = int(model.geom_type[g1]), int(model.geom_type[g2])
        out.append(2 if (t1 == GEOM.PLANE and t2 == GEOM.CAPSULE) else 1)
    return out


def build_model_arrays(model: Model, frictionloss_override: bool = False) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (header[MH_SIZE] int32, mi int32, mf float32) -- the header's task sizes are filled by pack_blob."""
    a = model.arrays
    nb, nv, nj = model.nbody, model.nv, model.njnt
    if nv > C.B2S_MAXNV or nb > C.B2S_MAXNB:
        raise ValueError(f"model too large for one dof/body per lane: nv={nv}, nbody={nb}")
    parent = a["body_parentid"]
    # depth-first ordering check: subtree(b) must be the contiguous range [b, end)
    sub_end = np.arange(1, nb + 1, dtype=np.int32)
    for b in range(nb - 1, 0, -1):
        sub_end[parent[b]] = max(sub_end[parent[b]], sub_end[b])
    sub_end[0] = nb
    for b in range(1, nb):
        for c in range(b + 1, sub_end[b]):
            p = c
            while p != b and p != 0:
                p = parent[p]
            if p != b:
                raise ValueError("bodies are not in depth-first order")
    depth = np.zeros(nb, dtype=np.int32)
    for b in range(1, nb):
        depth[b] = depth[parent[b]] + 1
    nlevel = int(depth.max()) if nb > 1 else 0
    level_body = np.array(sorted(range(1, nb), key=lambda b: (depth[b], b)), dtype=np.int32)
    level_start = np.zeros(nlevel + 1, dtype=np.int32)
    for lvl in range(1, nlevel + 1):
        level_start[lvl] = level_start[lvl - 1] + int((depth[1:] == lvl).sum())
    assert level_start[-1] == nb - 1

    # leaf-to-root accumulation schedule (composite inertia, subtree force sums): deepest level first, one pass per
    # sibling rank so that no two bodies of a pass share a parent; children of the world are never accumulated
    up_start, up_body = [0], []
    for lvl in range(nlevel, 1, -1):
        by_parent: dict[int, list[int]] = {}
        for b in range(1, nb):
            if depth[b] == lvl:
                by_parent.setdefault(int(parent[b]), []).append(b)
        for rank in range(max((len(v) for v in by_parent.values()), default=0)):
            up_body += [v[rank] for v in by_parent.values() if len(v) > rank]
            up_start.append(len(up_body))
    up_start_a, up_body_a = np.array(up_start, dtype=np.int32), np.array(up_body, dtype=np.int32)
    up_child = np.full((len(up_start) - 1, nb), -1, dtype=np.int32)
    for pi in range(len(up_start) - 1):
        for b in up_body[up_start[pi] : up_start[pi + 1]]:
            assert up_child[pi, parent[b]] == -1
            up_child[pi, parent[b]] = b

    dof_act = np.full(nv, -1, dtype=np.int32)
    for u, jid in enumerate(a["actuator_trnid"]):
        da = int(a["jnt_dofadr"][int(jid)])
        dof_act[da] = u if dof_act[da] == -1 else -2

    dof_parent = a["dof_parentid"]
    body_lastdof = np.full(nb, -1, dtype=np.int32)
    for b in range(1, nb):
        if a["body_dofnum"][b] > 0:
            body_lastdof[b] = a["body_dofadr"][b] + a["body_dofnum"][b] - 1
        else:
            body_lastdof[b] = body_lastdof[parent[b]]
    body_dofmask = np.zeros(nb, dtype=np.int64)
    for b in range(1, nb):
        j = int(body_lastdof[b])
        while j >= 0:
            body_dofmask[b] |= 1 << j
            j = int(dof_parent[j])
    dof_velparent = dof_parent.copy().astype(np.int32)
    for j in range(nj):
        if a["jnt_type"][j] == JNT.FREE:
            da = int(a["jnt_dofadr"][j])
            dof_velparent[da : da + 3] = -2
            dof_velparent[da + 3 : da + 6] = da + 2
    limit_jnt = np.array([j for j in range(nj) if a["jnt_type"][j] == JNT.HINGE and a["jnt_limited"][j]], dtype=np.int32)
    ncon_pair = contacts_per_pair(model)
    conadr = np.concatenate([[0], np.cumsum(ncon_pair)[:-1]]).astype(np.int32) if ncon_pair else np.zeros(0, np.int32)
    ncon = int(sum(ncon_pair))
    ndense = int(sum(n * (1 if d == 1 else 4) for n, d in zip(ncon_pair, a["pair_dim"])))
    nfric = int((a["dof_frictionloss"] > 0).sum())
    if frictionloss_override:
        nfric = nv

    ints: list[tuple[int, np.ndarray]] = [
        (C.MH_BODY_PARENTID, parent), (C.MH_BODY_ROOTID, a["body_rootid"]), (C.MH_BODY_JNTNUM, a["body_jntnum"]),
        (C.MH_BODY_JNTADR, a["body_jntadr"]), (C.MH_BODY_DOFNUM, a["body_dofnum"]), (C.MH_BODY_DOFADR, a["body_dofadr"]),
        (C.MH_LEVEL_START, level_start), (C.MH_LEVEL_BODY, level_body),
        (C.MH_JNT_TYPE, a["jnt_type"]), (C.MH_JNT_QPOSADR, a["jnt_qposadr"]), (C.MH_JNT_DOFADR, a["jnt_dofadr"]),
        (C.MH_JNT_BODYID, a["jnt_bodyid"]), (C.MH_JNT_LIMITED, a["jnt_limited"]),
        (C.MH_DOF_BODYID, a["dof_bodyid"]), (C.MH_DOF_JNTID, a["dof_jntid"]), (C.MH_DOF_PARENTID, dof_parent),
        (C.MH_GEOM_TYPE, a["geom_type"]), (C.MH_GEOM_BODYID, a["geom_bodyid"]), (C.MH_BODY_SUBTREE_END, sub_end),
        (C.MH_SITE_BODYID, a["site_bodyid"]), (C.MH_SITE_TYPE, a["site_type"]),
        (C.MH_PAIR_GEOM1, a["pair_geom1"]), (C.MH_PAIR_GEOM2, a["pair_geom2"]), (C.MH_PAIR_DIM, a["pair_dim"]),
        (C.MH_ACT_TRNID, a["actuator_trnid"]), (C.MH_ACT_CTRLLIMITED, a["actuator_ctrllimited"]),
        (C.MH_ACT_FORCELIMITED, a["actuator_forcelimited"]),
        (C.MH_SENSOR_TYPE, a["sensor_type"]), (C.MH_SENSOR_OBJTYPE, a["sensor_objtype"]), (C.MH_SENSOR_OBJID, a["sensor_objid"]),
        (C.MH_SENSOR_DIM, a["sensor_dim"]), (C.MH_SENSOR_ADR, a["sensor_adr"]),
        (C.MH_LIMIT_JNT, limit_jnt), (C.MH_PAIR_CONADR, conadr),
        (C.MH_BODY_DOFMASK, body_dofmask.astype(np.uint32).view(np.int32)),
        (C.MH_DOF_VELPARENT, dof_velparent), (C.MH_BODY_LASTDOF, body_lastdof),
        (C.MH_UP_START, up_start_a), (C.MH_UP_BODY, up_body_a), (C.MH_DOF_ACT, dof_act),
        (C.MH_UP_CHILD, up_child.reshape(-1)),
    ]
    o = model.opt
    scalars = np.zeros(C.MF_SCALARS, dtype=np.float32)
    scalars[C.MF_TIMESTEP] = o.timestep
    scalars[C.MF_GRAVITY : C.MF_GRAVITY + 3] = o.gravity
    scalars[C.MF_TOLERANCE] = o.tolerance
    scalars[C.MF_LS_TOLERANCE] = o.ls_tolerance
    scalars[C.MF_IMPRATIO] = o.impratio
    scalars[C.MF_MEANINERTIA] = model.meaninertia
    floats: list[tuple[int, np.ndarray]] = [
        (C.MH_BODY_POS, a["body_pos"]), (C.MH_BODY_QUAT, a["body_quat"]), (C.MH_BODY_IPOS, a["body_ipos"]),
        (C.MH_BODY_IQUAT, a["body_iquat"]), (C.MH_BODY_MASS, a["body_mass"]), (C.MH_BODY_INERTIA, a["body_inertia"]),
        (C.MH_BODY_SUBTREEMASS, a["body_subtreemass"]), (C.MH_BODY_INVWEIGHT0, a["body_invweight0"]),
        (C.MH_JNT_POS, a["jnt_pos"]), (C.MH_JNT_AXIS, a["jnt_axis"]), (C.MH_JNT_STIFFNESS, a["jnt_stiffness"]),
        (C.MH_JNT_RANGE, a["jnt_range"]), (C.MH_JNT_MARGIN, a["jnt_margin"]), (C.MH_JNT_SOLREF, a["jnt_solref"]),
        (C.MH_JNT_SOLIMP, a["jnt_solimp"]),
        (C.MH_DOF_ARMATURE, a["dof_armature"]), (C.MH_DOF_DAMPING, a["dof_damping"]),
        (C.MH_DOF_FRICTIONLOSS, a["dof_frictionloss"]), (C.MH_DOF_INVWEIGHT0, a["dof_invweight0"]),
        (C.MH_DOF_SOLREF, a["dof_solref"]), (C.MH_DOF_SOLIMP, a["dof_solimp"]),
        (C.MH_QPOS0, a["qpos0"]), (C.MH_QPOS_SPRING, a["qpos_spring"]),
        (C.MH_GEOM_POS, a["geom_pos"]), (C.MH_GEOM_QUAT, a["geom_quat"]), (C.MH_GEOM_SIZE, a["geom_size"]),
        (C.MH_SITE_POS, a["site_pos"]), (C.MH_SITE_QUAT, a["site_quat"]), (C.MH_SITE_SIZE, a["site_size"]),
        (C.MH_PAIR_FRICTION, a["pair_friction"]), (C.MH_PAIR_SOLREF, a["pair_solref"]), (C.MH_PAIR_SOLIMP, a["pair_solimp"]),
        (C.MH_PAIR_MARGIN, a["pair_margin"]), (C.MH_PAIR_GAP, a["pair_gap"]),
        (C.MH_ACT_GEAR, a["actuator_gear"]), (C.MH_ACT_CTRLRANGE, a["actuator_ctrlrange"]),
        (C.MH_ACT_FORCERANGE, a["actuator_forcerange"]),
    ]
    header = np.zeros(C.MH_SIZE, dtype=np.int32)
    mi_parts, off = [], 0
    for slot, arr in ints:
        flat = np.ascontiguousarray(arr, dtype=np.int32).ravel()
        header[slot] = off
        mi_parts.append(flat)
        off += flat.size
    mi = np.concatenate(mi_parts) if mi_parts else np.zeros(0, np.int32)
    mf_parts, off = [scalars], scalars.size
    for slot, arr in floats:
        flat = np.ascontiguousarray(arr, dtype=np.float64).ravel().astype(np.float32)
        header[slot] = off
        mf_parts.append(flat)
        off += flat.size
    mf = np.concatenate(mf_parts)
    header[C.MH_MAGIC] = C.B2S_MODEL_MAGIC
    header[C.MH_VERSION] = C.B2S_MODEL_VERSION
    header[C.MH_NI], header[C.MH_NF] = mi.size, mf.size
    header[C.MH_NQ], header[C.MH_NV], header[C.MH_NU] = model.nq, nv, model.nu
    header[C.MH_NBODY], header[C.MH_NJNT], header[C.MH_NGEOM] = nb, nj, model.ngeom
    header[C.MH_NSITE], header[C.MH_NSENSOR], header[C.MH_NSENSORDATA] = model.nsite, model.nsensor, model.nsensordata
    header[C.MH_NPAIR], header[C.MH_NCON], header[C.MH_NEFC_DENSE] = model.npair, ncon, ndense
    header[C.MH_NLEVEL] = nlevel
    header[C.MH_ITERATIONS], header[C.MH_LS_ITERATIONS] = o.iterations, o.ls_iterations
    header[C.MH_DISABLE_REFSAFE], header[C.MH_DISABLE_WARMSTART] = int(o.disable_refsafe), int(o.disable_warmstart)
    header[C.MH_NFRICTION], header[C.MH_NLIMIT] = nfric, limit_jnt.size
    header[C.MH_NUPPASS] = len(up_start) - 1
    return header, mi, mf


def pack_blob(model: Model, ti: np.ndarray, tf: np.ndarray, frictionloss_override: bool = False) -> bytes:
    header, mi, mf = build_model_arrays(model, frictionloss_override)
    ti = np.ascontiguousarray(ti, dtype=np.int32)
    tf = np.ascontiguousarray(tf, dtype=np.float32)
    header[C.MH_NTI], header[C.MH_NTF] = ti.size, tf.size
    return header.tobytes() + mi.tobytes() + mf.tobytes() + ti.tobytes() + tf.tobytes()
