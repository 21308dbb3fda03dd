/*
 * Oracle model I/O: fills an OModel field-by-field from flat double arrays handed over by Python
 * (ksim_b200.mjcf.Model.arrays).  TEST INFRASTRUCTURE -- see oracle.h.
 */
#include "oracle.h"

#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct { const char* name; size_t offset; int count; int is_int; } Field;

#define FR(f, n) {#f, offsetof(OModel, f), (n), 0}
#define FI(f, n) {#f, offsetof(OModel, f), (n), 1}

static const Field FIELDS[] = {
    FI(nq, 1), FI(nv, 1), FI(nu, 1), FI(nbody, 1), FI(njnt, 1), FI(ngeom, 1), FI(nsite, 1), FI(nsensor, 1),
    FI(nsensordata, 1), FI(npair, 1),
    FR(timestep, 1), FR(gravity, 3), FR(tolerance, 1), FR(ls_tolerance, 1), FR(impratio, 1), FR(meaninertia, 1),
    FI(iterations, 1), FI(ls_iterations, 1), FI(disable_refsafe, 1), FI(disable_warmstart, 1),
    FI(body_parentid, O_MAXNB), FI(body_rootid, O_MAXNB), FI(body_jntnum, O_MAXNB), FI(body_jntadr, O_MAXNB),
    FI(body_dofnum, O_MAXNB), FI(body_dofadr, O_MAXNB),
    FR(body_pos, O_MAXNB * 3), FR(body_quat, O_MAXNB * 4), FR(body_ipos, O_MAXNB * 3), FR(body_iquat, O_MAXNB * 4),
    FR(body_mass, O_MAXNB), FR(body_inertia, O_MAXNB * 3), FR(body_subtreemass, O_MAXNB),
    FR(body_invweight0, O_MAXNB * 2),
    FI(jnt_type, O_MAXNJ), FI(jnt_qposadr, O_MAXNJ), FI(jnt_dofadr, O_MAXNJ), FI(jnt_bodyid, O_MAXNJ),
    FI(jnt_limited, O_MAXNJ),
    FR(jnt_pos, O_MAXNJ * 3), FR(jnt_axis, O_MAXNJ * 3), FR(jnt_stiffness, O_MAXNJ), FR(jnt_range, O_MAXNJ * 2),
    FR(jnt_margin, O_MAXNJ), FR(jnt_solref, O_MAXNJ * 2), FR(jnt_solimp, O_MAXNJ * 5),
    FI(dof_bodyid, O_MAXNV), FI(dof_jntid, O_MAXNV), FI(dof_parentid, O_MAXNV),
    FR(dof_armature, O_MAXNV), FR(dof_damping, O_MAXNV), FR(dof_frictionloss, O_MAXNV), FR(dof_invweight0, O_MAXNV),
    FR(dof_solref, O_MAXNV * 2), FR(dof_solimp, O_MAXNV * 5),
    FR(qpos0, O_MAXNQ), FR(qpos_spring, O_MAXNQ),
    FI(geom_type, O_MAXNG), FI(geom_bodyid, O_MAXNG),
    FR(geom_pos, O_MAXNG * 3), FR(geom_quat, O_MAXNG * 4), FR(geom_size, O_MAXNG * 3),
    FI(site_bodyid, O_MAXNS), FR(site_pos, O_MAXNS * 3), FR(site_quat, O_MAXNS * 4), FR(site_size, O_MAXNS * 3),
    FI(site_type, O_MAXNS),
    FI(pair_geom1, O_MAXPAIR), FI(pair_geom2, O_MAXPAIR), FI(pair_dim, O_MAXPAIR),
    FR(pair_friction, O_MAXPAIR * 5), FR(pair_solref, O_MAXPAIR * 2), FR(pair_solimp, O_MAXPAIR * 5),
    FR(pair_margin, O_MAXPAIR), FR(pair_gap, O_MAXPAIR),
    FI(actuator_trnid, O_MAXNU), FI(actuator_ctrllimited, O_MAXNU), FI(actuator_forcelimited, O_MAXNU),
    FR(actuator_gear, O_MAXNU), FR(actuator_ctrlrange, O_MAXNU * 2), FR(actuator_forcerange, O_MAXNU * 2),
    FI(sensor_type, O_MAXSENS), FI(sensor_objtype, O_MAXSENS), FI(sensor_objid, O_MAXSENS),
    FI(sensor_dim, O_MAXSENS), FI(sensor_adr, O_MAXSENS),
};

OModel* o_model_new(void) { return (OModel*)calloc(1, sizeof(OModel)); }
void o_model_free(OModel* m) { free(m); }
OModel* o_model_clone(const OModel* m) {
  OModel* c = (OModel*)malloc(sizeof(OModel));
  memcpy(c, m, sizeof(OModel));
  return c;
}

/* returns 0 on success, 1 unknown field, 2 too many values */
int o_model_set(OModel* m, const char* name, const double* data, int n) {
  for (size_t f = 0; f < sizeof(FIELDS) / sizeof(FIELDS[0]); f++) {
    if (strcmp(FIELDS[f].name, name) != 0) continue;
    if (n > FIELDS[f].count) return 2;
    char* base = (char*)m + FIELDS[f].offset;
    if (FIELDS[f].is_int) { int* p = (int*)base; for (int i = 0; i < n; i++) p[i] = (int)data[i]; }
    else { REAL* p = (REAL*)base; for (int i = 0; i < n; i++) p[i] = (REAL)data[i]; }
    return 0;
  }
  return 1;
}

int o_model_get(const OModel* m, const char* name, double* data, int n) {
  for (size_t f = 0; f < sizeof(FIELDS) / sizeof(FIELDS[0]); f++) {
    if (strcmp(FIELDS[f].name, name) != 0) continue;
    if (n > FIELDS[f].count) return 2;
    const char* base = (const char*)m + FIELDS[f].offset;
    if (FIELDS[f].is_int) { const int* p = (const int*)base; for (int i = 0; i < n; i++) data[i] = p[i]; }
    else { const REAL* p = (const REAL*)base; for (int i = 0; i < n; i++) data[i] = (double)p[i]; }
    return 0;
  }
  return 1;
}

/* ---- OData access by name for tests ---- */
typedef struct { const char* name; size_t offset; int count; int is_int; } DField;
#define DR(f, n) {#f, offsetof(OData, f), (n), 0}
#define DI(f, n) {#f, offsetof(OData, f), (n), 1}
static const DField DFIELDS[] = {
    DR(qpos, O_MAXNQ), DR(qvel, O_MAXNV), DR(qacc, O_MAXNV), DR(qacc_warmstart, O_MAXNV), DR(ctrl, O_MAXNU),
    DR(time, 1), DR(xfrc_applied, O_MAXNB * 6),
    DR(xpos, O_MAXNB * 3), DR(xquat, O_MAXNB * 4), DR(xmat, O_MAXNB * 9), DR(xipos, O_MAXNB * 3),
    DR(ximat, O_MAXNB * 9), DR(xanchor, O_MAXNJ * 3), DR(xaxis, O_MAXNJ * 3),
    DR(geom_xpos, O_MAXNG * 3), DR(geom_xmat, O_MAXNG * 9), DR(site_xpos, O_MAXNS * 3), DR(site_xmat, O_MAXNS * 9),
    DR(subtree_com, O_MAXNB * 3), DR(cinert, O_MAXNB * 10), DR(crb, O_MAXNB * 10), DR(cdof, O_MAXNV * 6),
    DR(qM, O_MAXNV * O_MAXNV), DR(qL, O_MAXNV * O_MAXNV), DI(ncon, 1), DI(nefc, 1),
    DR(cvel, O_MAXNB * 6), DR(cdof_dot, O_MAXNV * 6),
    DR(qfrc_passive, O_MAXNV), DR(qfrc_bias, O_MAXNV), DR(qfrc_actuator, O_MAXNV), DR(actuator_force, O_MAXNU),
    DR(qfrc_smooth, O_MAXNV), DR(qacc_smooth, O_MAXNV), DR(qfrc_constraint, O_MAXNV),
    DI(efc_type, O_MAXEFC), DI(efc_id, O_MAXEFC), DR(efc_J, O_MAXEFC * O_MAXNV), DR(efc_pos, O_MAXEFC),
    DR(efc_aref, O_MAXEFC), DR(efc_D, O_MAXEFC), DR(efc_R, O_MAXEFC), DR(efc_force, O_MAXEFC),
    DI(solver_niter, 1), DR(cacc, O_MAXNB * 6), DR(cfrc_int, O_MAXNB * 6), DR(cfrc_ext, O_MAXNB * 6),
    DR(sensordata, O_MAXSD),
};

OData* o_data_new(const OModel* m) { OData* d = (OData*)malloc(sizeof(OData)); o_make_data(m, d); return d; }
void o_data_free(OData* d) { free(d); }

static const DField* dfind(const char* name) {
  for (size_t f = 0; f < sizeof(DFIELDS) / sizeof(DFIELDS[0]); f++)
    if (strcmp(DFIELDS[f].name, name) == 0) return &DFIELDS[f];
  return 0;
}
int o_data_set(OData* d, const char* name, const double* data, int n) {
  const DField* f = dfind(name);
  if (!f) return 1;
  if (n > f->count) return 2;
  char* base = (char*)d + f->offset;
  if (f->is_int) { int* p = (int*)base; for (int i = 0; i < n; i++) p[i] = (int)data[i]; }
  else { REAL* p = (REAL*)base; for (int i = 0; i < n; i++) p[i] = (REAL)data[i]; }
  return 0;
}
int o_data_get(const OData* d, const char* name, double* data, int n) {
  const DField* f = dfind(name);
  if (!f) return 1;
  if (n > f->count) return 2;
  const char* base = (const char*)d + f->offset;
  if (f->is_int) { const int* p = (const int*)base; for (int i = 0; i < n; i++) data[i] = p[i]; }
  else { const REAL* p = (const REAL*)base; for (int i = 0; i < n; i++) data[i] = (double)p[i]; }
  return 0;
}
int o_contact_get(const OData* d, int c, double* dist, double* pos, int* geom) {
  if (c < 0 || c >= d->ncon) return 1;
  *dist = d->contact[c].dist;
  for (int k = 0; k < 3; k++) pos[k] = d->contact[c].pos[k];
  geom[0] = d->contact[c].geom1; geom[1] = d->contact[c].geom2;
  return 0;
}
int o_sizeof_real(void) { return (int)sizeof(REAL); }
