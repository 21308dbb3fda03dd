/*
 * b200sim model blob layout.  The host (ksim_b200/blob.py) flattens a compiled rigid-body model
 * (the fields `mjx.put_model` would carry, ksim/utils/mujoco.py:388-391) plus the task program
 * (include/b200sim_codes.h) into one byte blob; b200sim_model_create() uploads it.
 *
 *   blob = int32 header[MH_SIZE] | int32 mi[header[MH_NI]] | float32 mf[header[MH_NF]]
 *        | int32 ti[header[MH_NTI]] | float32 tf[header[MH_NTF]]
 *
 * header slots below MH_FIRST_OFF are scalars; from MH_FIRST_OFF on they are offsets into mi[] (MI_*) or
 * mf[] (MF_*).  Python parses the #defines of this file (ksim_b200/codes.py); keep it plain.
 */
#ifndef B200SIM_MODEL_H
#define B200SIM_MODEL_H

#define B2S_MODEL_MAGIC 0x4232534D /* "B2SM" */
#define B2S_MODEL_VERSION 6

#define B2S_MAXNV 32   /* one dof per lane */
#define B2S_MAXNB 32   /* one body per lane */
#define B2S_MAXNU 32
#define B2S_MAXPAIR 8
#define B2S_MAXCON 16
#define B2S_MAXLEVEL 16

/* ---- header scalars ---- */
#define MH_MAGIC 0
#define MH_VERSION 1
#define MH_NI 2
#define MH_NF 3
#define MH_NTI 4
#define MH_NTF 5
#define MH_NQ 6
#define MH_NV 7
#define MH_NU 8
#define MH_NBODY 9
#define MH_NJNT 10
#define MH_NGEOM 11
#define MH_NSITE 12
#define MH_NSENSOR 13
#define MH_NSENSORDATA 14
#define MH_NPAIR 15
#define MH_NCON 16     /* candidate contacts per step (plane-sphere 1, plane-capsule 2, sphere-sphere 1 per pair) */
#define MH_NEFC_DENSE 17 /* max dense (contact) constraint rows */
#define MH_NLEVEL 18   /* kinematic tree depth (levels of bodies below the world) */
#define MH_ITERATIONS 19
#define MH_LS_ITERATIONS 20
#define MH_DISABLE_REFSAFE 21
#define MH_DISABLE_WARMSTART 22
#define MH_NFRICTION 23 /* dofs with base frictionloss > 0 or a frictionloss override (rows are emitted per env) */
#define MH_NLIMIT 24    /* limited hinge joints */
#define MH_NUPPASS 25   /* passes of the leaf-to-root accumulation schedule (MH_UP_START / MH_UP_BODY) */
/* float scalars live at the start of mf[] */
#define MF_TIMESTEP 0
#define MF_GRAVITY 1   /* 3 */
#define MF_TOLERANCE 4
#define MF_LS_TOLERANCE 5
#define MF_IMPRATIO 6
#define MF_MEANINERTIA 7
#define MF_SCALARS 8

/* ---- header offsets into mi[] ---- */
#define MH_FIRST_OFF 32
#define MH_BODY_PARENTID 32
#define MH_BODY_ROOTID 33
#define MH_BODY_JNTNUM 34
#define MH_BODY_JNTADR 35
#define MH_BODY_DOFNUM 36
#define MH_BODY_DOFADR 37
#define MH_LEVEL_START 38   /* [nlevel+1] into LEVEL_BODY */
#define MH_LEVEL_BODY 39    /* bodies 1..nbody-1 sorted by depth */
#define MH_JNT_TYPE 40
#define MH_JNT_QPOSADR 41
#define MH_JNT_DOFADR 42
#define MH_JNT_BODYID 43
#define MH_JNT_LIMITED 44
#define MH_DOF_BODYID 45
#define MH_DOF_JNTID 46
#define MH_DOF_PARENTID 47
#define MH_GEOM_TYPE 48
#define MH_GEOM_BODYID 49
#define MH_BODY_SUBTREE_END 50 /* [nbody] bodies are depth-first ordered: subtree(b) = [b, end) */
#define MH_SITE_BODYID 51
#define MH_SITE_TYPE 52
#define MH_PAIR_GEOM1 53
#define MH_PAIR_GEOM2 54
#define MH_PAIR_DIM 55
#define MH_ACT_TRNID 56     /* joint id */
#define MH_ACT_CTRLLIMITED 57
#define MH_ACT_FORCELIMITED 58
#define MH_SENSOR_TYPE 59
#define MH_SENSOR_OBJTYPE 60
#define MH_SENSOR_OBJID 61
#define MH_SENSOR_DIM 62
#define MH_SENSOR_ADR 63
#define MH_LIMIT_JNT 64     /* [nlimit] joint ids of limited hinges, in joint order */
#define MH_PAIR_CONADR 65    /* [npair] first candidate-contact slot of the pair */
#define MH_BODY_DOFMASK 66   /* [nbody] bitmask of dofs that move the body */
#define MH_DOF_VELPARENT 67  /* [nv] dof whose chain gives the velocity "before" this dof; -2 = none (free translation) */
#define MH_BODY_LASTDOF 68   /* [nbody] last dof of the body or of its nearest ancestor with dofs; -1 for the world */
/* Leaf-to-root accumulation schedule: pass p adds bodies MH_UP_BODY[MH_UP_START[p] .. MH_UP_START[p+1]) into their
 * parents.  Passes go from the deepest level up; within a level, pass r holds each parent's r-th child, so the bodies
 * of one pass have distinct parents and a pass is one conflict-free lane-parallel phase. */
#define MH_UP_START 69       /* [nuppass+1] */
#define MH_UP_BODY 70        /* [nbody-1 minus the children of the world] */
#define MH_UP_CHILD 72       /* [nuppass][nbody] the child that pass p adds into body b, or -1 (inverse of MH_UP_BODY) */
#define MH_DOF_ACT 71        /* [nv] the actuator that drives the dof; -1 none; -2 several (the kernel gathers) */
/* ---- header offsets into mf[] ---- */
#define MH_BODY_POS 80
#define MH_BODY_QUAT 81
#define MH_BODY_IPOS 82
#define MH_BODY_IQUAT 83
#define MH_BODY_MASS 84
#define MH_BODY_INERTIA 85
#define MH_BODY_SUBTREEMASS 86
#define MH_BODY_INVWEIGHT0 87   /* [nbody][2] */
#define MH_JNT_POS 88
#define MH_JNT_AXIS 89
#define MH_JNT_STIFFNESS 90
#define MH_JNT_RANGE 91
#define MH_JNT_MARGIN 92
#define MH_JNT_SOLREF 93
#define MH_JNT_SOLIMP 94
#define MH_DOF_ARMATURE 95
#define MH_DOF_DAMPING 96
#define MH_DOF_FRICTIONLOSS 97
#define MH_DOF_INVWEIGHT0 98
#define MH_DOF_SOLREF 99
#define MH_DOF_SOLIMP 100
#define MH_QPOS0 101
#define MH_QPOS_SPRING 102
#define MH_GEOM_POS 103
#define MH_GEOM_QUAT 104
#define MH_GEOM_SIZE 105
#define MH_SITE_POS 106
#define MH_SITE_QUAT 107
#define MH_SITE_SIZE 108
#define MH_PAIR_FRICTION 109   /* [npair][5] */
#define MH_PAIR_SOLREF 110
#define MH_PAIR_SOLIMP 111
#define MH_PAIR_MARGIN 112
#define MH_PAIR_GAP 113
#define MH_ACT_GEAR 114
#define MH_ACT_CTRLRANGE 115
#define MH_ACT_FORCERANGE 116
#define MH_SIZE 128

#endif
