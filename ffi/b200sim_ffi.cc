// XLA FFI handlers over the C-ABI of include/b200sim.h: what a ksim maintainer links into the JAX process so that a
// `B200Engine(PhysicsEngine)` can call the CUDA path from inside `jit` / `vmap` / `scan`
// (`jax.ffi.ffi_call(target, ..., vmap_method="broadcast_all")`).  Written against the FFI *C* API (`xla/ffi/api/c_api.h`:
// one `XLA_FFI_Error* handler(XLA_FFI_CallFrame*)` per target), so it needs no C++ binding templates.
//
//   target                      reference call it replaces                       operands -> results
//   b200sim_ffi_engine_reset    PhysicsEngine.reset  (ksim/engine.py:205-246)    rng u32[N,2], level f32[N], rand f32[N,RAND]
//                                                                                  -> state f32[N,S], keys u32[N,K], data f32[N,DATA]
//   b200sim_ffi_engine_step     PhysicsEngine.step   (ksim/engine.py:259-361)    action f32[N,NA], rng u32[N,2], rand, state, keys
//                                                                                  -> state, keys, data
//   b200sim_ffi_step_ctrl       RLTask.step_engine   (ksim/task/rl.py:1164-1211) action, rand, state, keys, rec_cur f32[N,R]
//                                                                                  -> state, keys, rec_cur, rec_next
//   b200sim_ffi_rewards         get_rewards          (ksim/task/rl.py:189-245)   rec f32[T,N,R], level f32[N], carry f32[N,C]
//                                                                                  -> carry, total f32[N,T], comps f32[K,N,T]
//   b200sim_ffi_gae             compute_ppo_inputs   (ksim/task/ppo.py:54-121)   values, rewards f32[B,T], dones, successes u8[B,T]
//                                                                                  -> advantages, targets, returns f32[B,T]
//
// Attributes: `model` (u64 / s64 scalar: the b200sim_model* returned by b200sim_model_create, created on the host before
// tracing) for the first four; `gamma`, `lam` (f32) and `flags` (s32) for GAE.  State-like operands that are also results are
// updated in place when XLA aliases them (`input_output_aliases`, the reference donates `physics_state`, engine.py:261), and
// copied device-to-device on the call's stream otherwise.  No host synchronisation, no allocation, errors as XLA_FFI_Error.
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "xla/ffi/api/c_api.h"

#include "b200sim.h"

namespace {

XLA_FFI_Error* make_error(const XLA_FFI_Api* api, XLA_FFI_Error_Code code, const std::string& msg) {
  XLA_FFI_Error_Create_Args a;
  memset(&a, 0, sizeof(a));
  a.struct_size = sizeof(a);
  a.message = msg.c_str();
  a.errc = code;
  return api->XLA_FFI_Error_Create(&a);
}

struct Call {
  XLA_FFI_CallFrame* f;
  std::string err;  // the first problem found

  void note(const std::string& msg) { if (err.empty()) err = msg; }

  bool execute_stage() const { return f->stage == XLA_FFI_ExecutionStage_EXECUTE; }

  const XLA_FFI_Buffer* buffer(bool ret, int64_t i, XLA_FFI_DataType dtype, const char* name) {
    const int64_t n = ret ? f->rets.size : f->args.size;
    if (i >= n) { note(std::string(ret ? "missing result " : "missing operand ") + name); return nullptr; }
    if (ret ? f->rets.types[i] != XLA_FFI_RetType_BUFFER : f->args.types[i] != XLA_FFI_ArgType_BUFFER) { note(std::string(name) + " is not a buffer"); return nullptr; }
    const XLA_FFI_Buffer* b = static_cast<const XLA_FFI_Buffer*>(ret ? f->rets.rets[i] : f->args.args[i]);
    if (b->struct_size < sizeof(XLA_FFI_Buffer)) { note("XLA_FFI_Buffer is smaller than the header this shim was built against"); return nullptr; }
    if (b->dtype != dtype) { note(std::string(name) + " has the wrong element type"); return nullptr; }
    return b;
  }
  const XLA_FFI_Buffer* arg(int64_t i, XLA_FFI_DataType t, const char* name) { return buffer(false, i, t, name); }
  const XLA_FFI_Buffer* ret(int64_t i, XLA_FFI_DataType t, const char* name) { return buffer(true, i, t, name); }

  const XLA_FFI_Scalar* scalar_attr(const char* name) {
    for (int64_t i = 0; i < f->attrs.size; i++) {
      const XLA_FFI_ByteSpan* s = f->attrs.names[i];
      if (s->len == strlen(name) && memcmp(s->ptr, name, s->len) == 0) {
        if (f->attrs.types[i] != XLA_FFI_AttrType_SCALAR) { note(std::string("attribute ") + name + " is not a scalar"); return nullptr; }
        return static_cast<const XLA_FFI_Scalar*>(f->attrs.attrs[i]);
      }
    }
    note(std::string("missing attribute ") + name);
    return nullptr;
  }
  const b200sim_model* model() {
    const XLA_FFI_Scalar* s = scalar_attr("model");
    if (!s) return nullptr;
    if (s->dtype != XLA_FFI_DataType_U64 && s->dtype != XLA_FFI_DataType_S64) { note("attribute model must be a 64-bit integer (the handle)"); return nullptr; }
    const b200sim_model* m = reinterpret_cast<const b200sim_model*>(static_cast<uintptr_t>(*static_cast<const uint64_t*>(s->value)));
    if (!m) note("attribute model is null");
    return m;
  }
  bool f32_attr(const char* name, float* out) {
    const XLA_FFI_Scalar* s = scalar_attr(name);
    if (!s) return false;
    if (s->dtype == XLA_FFI_DataType_F32) *out = *static_cast<const float*>(s->value);
    else if (s->dtype == XLA_FFI_DataType_F64) *out = (float)*static_cast<const double*>(s->value);
    else { note(std::string("attribute ") + name + " must be a float"); return false; }
    return true;
  }
  bool i32_attr(const char* name, int* out) {
    const XLA_FFI_Scalar* s = scalar_attr(name);
    if (!s) return false;
    if (s->dtype == XLA_FFI_DataType_S32) *out = *static_cast<const int32_t*>(s->value);
    else if (s->dtype == XLA_FFI_DataType_S64) *out = (int)*static_cast<const int64_t*>(s->value);
    else { note(std::string("attribute ") + name + " must be an integer"); return false; }
    return true;
  }
  void* stream() {
    XLA_FFI_Stream_Get_Args a;
    memset(&a, 0, sizeof(a));
    a.struct_size = sizeof(a);
    a.ctx = f->ctx;
    if (XLA_FFI_Error* e = f->api->XLA_FFI_Stream_Get(&a)) { (void)e; note("XLA_FFI_Stream_Get failed"); return nullptr; }
    return a.stream;
  }
};

size_t count(const XLA_FFI_Buffer* b) {
  size_t n = 1;
  for (int64_t i = 0; i < b->rank; i++) n *= (size_t)b->dims[i];
  return n;
}

// in/out operand: in place when XLA aliased result to operand, one device-to-device copy on the stream otherwise
bool carry_over(const XLA_FFI_Buffer* in, const XLA_FFI_Buffer* out, size_t elem, void* stream, std::string* err) {
  if (count(in) != count(out)) { *err = "in/out operand and result differ in size"; return false; }
  if (in->data == out->data) return true;
  if (cudaMemcpyAsync(out->data, in->data, count(in) * elem, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) {
    *err = "device-to-device copy of an in/out operand failed";
    return false;
  }
  return true;
}

#define FAIL_IF(cond) do { if (cond) return make_error(frame->api, XLA_FFI_Error_Code_INVALID_ARGUMENT, c.err.empty() ? #cond : c.err); } while (0)
#define CHECK_RC(rc, errfn) do { if ((rc) != 0) return make_error(frame->api, XLA_FFI_Error_Code_INTERNAL, std::string(errfn())); } while (0)
#define PROLOGUE()                                                                  \
  if (frame->struct_size < sizeof(XLA_FFI_CallFrame)) return nullptr; /* a newer, smaller-prefix caller: nothing safe to do */ \
  Call c{frame, {}};                                                                \
  if (!c.execute_stage()) return nullptr /* nothing to instantiate / prepare / initialise */

}  // namespace

extern "C" {

int b200sim_ffi_built_against_stub(void) {
#ifdef B200SIM_XLA_FFI_STUB
  return 1;
#else
  return 0;
#endif
}

XLA_FFI_Error* b200sim_ffi_engine_reset(XLA_FFI_CallFrame* frame) {
  PROLOGUE();
  const b200sim_model* m = c.model();
  const XLA_FFI_Buffer *rng = c.arg(0, XLA_FFI_DataType_U32, "rng"), *level = c.arg(1, XLA_FFI_DataType_F32, "level"),
                       *rand = c.arg(2, XLA_FFI_DataType_F32, "rand");
  const XLA_FFI_Buffer *state = c.ret(0, XLA_FFI_DataType_F32, "state"), *keys = c.ret(1, XLA_FFI_DataType_U32, "keys"),
                       *data = c.ret(2, XLA_FFI_DataType_F32, "data");
  FAIL_IF(!m || !rng || !level || !rand || !state || !keys || !data);
  void* s = c.stream();
  FAIL_IF(!s && !c.err.empty());
  const int n = (int)count(level);
  CHECK_RC(b200sim_engine_reset(m, s, n, (const uint32_t*)rng->data, (const float*)level->data, (const float*)rand->data, (float*)state->data,
                                (uint32_t*)keys->data, (float*)data->data), b200sim_engine_last_error);
  return nullptr;
}

XLA_FFI_Error* b200sim_ffi_engine_step(XLA_FFI_CallFrame* frame) {
  PROLOGUE();
  const b200sim_model* m = c.model();
  const XLA_FFI_Buffer *action = c.arg(0, XLA_FFI_DataType_F32, "action"), *rng = c.arg(1, XLA_FFI_DataType_U32, "rng"),
                       *rand = c.arg(2, XLA_FFI_DataType_F32, "rand"), *state = c.arg(3, XLA_FFI_DataType_F32, "state"),
                       *keys = c.arg(4, XLA_FFI_DataType_U32, "keys");
  const XLA_FFI_Buffer *state_o = c.ret(0, XLA_FFI_DataType_F32, "state"), *keys_o = c.ret(1, XLA_FFI_DataType_U32, "keys"),
                       *data = c.ret(2, XLA_FFI_DataType_F32, "data");
  FAIL_IF(!m || !action || !rng || !rand || !state || !keys || !state_o || !keys_o || !data);
  void* s = c.stream();
  FAIL_IF(!s && !c.err.empty());
  FAIL_IF(!carry_over(state, state_o, 4, s, &c.err) || !carry_over(keys, keys_o, 4, s, &c.err));
  const int n = (int)(count(rng) / 2);
  CHECK_RC(b200sim_engine_step(m, s, n, (const float*)action->data, (const uint32_t*)rng->data, (const float*)rand->data, (float*)state_o->data,
                               (uint32_t*)keys_o->data, (float*)data->data), b200sim_engine_last_error);
  return nullptr;
}

XLA_FFI_Error* b200sim_ffi_step_ctrl(XLA_FFI_CallFrame* frame) {
  PROLOGUE();
  const b200sim_model* m = c.model();
  const XLA_FFI_Buffer *action = c.arg(0, XLA_FFI_DataType_F32, "action"), *rand = c.arg(1, XLA_FFI_DataType_F32, "rand"),
                       *state = c.arg(2, XLA_FFI_DataType_F32, "state"), *keys = c.arg(3, XLA_FFI_DataType_U32, "keys"),
                       *rec_cur = c.arg(4, XLA_FFI_DataType_F32, "rec_cur");
  const XLA_FFI_Buffer *state_o = c.ret(0, XLA_FFI_DataType_F32, "state"), *keys_o = c.ret(1, XLA_FFI_DataType_U32, "keys"),
                       *rec_cur_o = c.ret(2, XLA_FFI_DataType_F32, "rec_cur"), *rec_next = c.ret(3, XLA_FFI_DataType_F32, "rec_next");
  FAIL_IF(!m || !action || !rand || !state || !keys || !rec_cur || !state_o || !keys_o || !rec_cur_o || !rec_next);
  void* s = c.stream();
  FAIL_IF(!s && !c.err.empty());
  FAIL_IF(!carry_over(state, state_o, 4, s, &c.err) || !carry_over(keys, keys_o, 4, s, &c.err) || !carry_over(rec_cur, rec_cur_o, 4, s, &c.err));
  const int na = b200sim_model_info(m, B2S_INFO_ACTION_SIZE);
  FAIL_IF(na <= 0);
  const int n = (int)(count(action) / (size_t)na);
  CHECK_RC(b200sim_step_ctrl(m, s, n, (const float*)action->data, (const float*)rand->data, (float*)state_o->data, (uint32_t*)keys_o->data,
                             (float*)rec_cur_o->data, (float*)rec_next->data, nullptr), b200sim_last_error);
  return nullptr;
}

XLA_FFI_Error* b200sim_ffi_rewards(XLA_FFI_CallFrame* frame) {
  PROLOGUE();
  const b200sim_model* m = c.model();
  const XLA_FFI_Buffer *rec = c.arg(0, XLA_FFI_DataType_F32, "rec"), *level = c.arg(1, XLA_FFI_DataType_F32, "level"),
                       *carry = c.arg(2, XLA_FFI_DataType_F32, "carry");
  const XLA_FFI_Buffer *carry_o = c.ret(0, XLA_FFI_DataType_F32, "carry"), *total = c.ret(1, XLA_FFI_DataType_F32, "total"),
                       *comps = c.ret(2, XLA_FFI_DataType_F32, "comps");
  FAIL_IF(!m || !rec || !level || !carry || !carry_o || !total || !comps);
  FAIL_IF(rec->rank != 3);
  void* s = c.stream();
  FAIL_IF(!s && !c.err.empty());
  FAIL_IF(!carry_over(carry, carry_o, 4, s, &c.err));
  CHECK_RC(b200sim_rewards(m, s, (int)rec->dims[1], (int)rec->dims[0], (const float*)rec->data, (const float*)level->data, (float*)carry_o->data,
                           (float*)total->data, (float*)comps->data), b200sim_last_error);
  return nullptr;
}

XLA_FFI_Error* b200sim_ffi_gae(XLA_FFI_CallFrame* frame) {
  PROLOGUE();
  const XLA_FFI_Buffer *values = c.arg(0, XLA_FFI_DataType_F32, "values"), *rewards = c.arg(1, XLA_FFI_DataType_F32, "rewards"),
                       *dones = c.arg(2, XLA_FFI_DataType_U8, "dones"), *succ = c.arg(3, XLA_FFI_DataType_U8, "successes");
  const XLA_FFI_Buffer *adv = c.ret(0, XLA_FFI_DataType_F32, "advantages"), *tgt = c.ret(1, XLA_FFI_DataType_F32, "value_targets"),
                       *ret = c.ret(2, XLA_FFI_DataType_F32, "returns");
  float gamma = 0.0f, lam = 0.0f;
  int flags = 0;
  FAIL_IF(!values || !rewards || !dones || !succ || !adv || !tgt || !ret || !c.f32_attr("gamma", &gamma) || !c.f32_attr("lam", &lam) ||
          !c.i32_attr("flags", &flags));
  FAIL_IF(values->rank != 2);
  void* s = c.stream();
  FAIL_IF(!s && !c.err.empty());
  CHECK_RC(b200sim_gae(s, (int)values->dims[0], (int)values->dims[1], (const float*)values->data, (const float*)rewards->data,
                       (const uint8_t*)dones->data, (const uint8_t*)succ->data, gamma, lam, flags, (float*)adv->data, (float*)tgt->data,
                       (float*)ret->data), b200sim_last_error);
  return nullptr;
}

}  // extern "C"
