/* MINIMAL STAND-IN for XLA's FFI C API header (`xla/ffi/api/c_api.h`, shipped inside jaxlib as
 * `jaxlib/include/xla/ffi/api/c_api.h`), used ONLY to compile-check ffi/b200sim_ffi.cc on machines without jaxlib (this
 * repository's build container and GPU boxes have no jax).  It declares the subset of types the shim touches, written from the
 * published description of the FFI call frame (https://openxla.org/xla/custom_call); the field order follows that header as
 * recalled and is NOT verified against a jaxlib release.  A real build MUST use the real header: ffi/Makefile prefers
 * `$(JAXLIB_INCLUDE)` whenever `python -c "import jaxlib"` works, and the shim checks `struct_size` of everything it reads. */
#ifndef XLA_FFI_API_C_API_H_
#define XLA_FFI_API_C_API_H_

#include <stddef.h>
#include <stdint.h>

#define B200SIM_XLA_FFI_STUB 1

#ifdef __cplusplus
extern "C" {
#endif

typedef struct XLA_FFI_Api XLA_FFI_Api;
typedef struct XLA_FFI_InternalApi XLA_FFI_InternalApi;
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_ExecutionContext XLA_FFI_ExecutionContext;
typedef struct XLA_FFI_Future XLA_FFI_Future;

typedef enum { XLA_FFI_Extension_Metadata = 1 } XLA_FFI_Extension_Type;
typedef struct XLA_FFI_Extension_Base {
  size_t struct_size;
  XLA_FFI_Extension_Type type;
  struct XLA_FFI_Extension_Base* next;
} XLA_FFI_Extension_Base;

typedef struct XLA_FFI_Api_Version {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int major_version;
  int minor_version;
} XLA_FFI_Api_Version;

typedef enum {
  XLA_FFI_Error_Code_OK = 0,
  XLA_FFI_Error_Code_CANCELLED = 1,
  XLA_FFI_Error_Code_UNKNOWN = 2,
  XLA_FFI_Error_Code_INVALID_ARGUMENT = 3,
  XLA_FFI_Error_Code_INTERNAL = 13
} XLA_FFI_Error_Code;

typedef struct XLA_FFI_Error_Create_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  const char* message;
  XLA_FFI_Error_Code errc;
} XLA_FFI_Error_Create_Args;
typedef XLA_FFI_Error* XLA_FFI_Error_Create(XLA_FFI_Error_Create_Args* args);

typedef enum {
  XLA_FFI_DataType_INVALID = 0,
  XLA_FFI_DataType_PRED = 1,
  XLA_FFI_DataType_S8 = 2,
  XLA_FFI_DataType_S16 = 3,
  XLA_FFI_DataType_S32 = 4,
  XLA_FFI_DataType_S64 = 5,
  XLA_FFI_DataType_U8 = 6,
  XLA_FFI_DataType_U16 = 7,
  XLA_FFI_DataType_U32 = 8,
  XLA_FFI_DataType_U64 = 9,
  XLA_FFI_DataType_F16 = 10,
  XLA_FFI_DataType_F32 = 11,
  XLA_FFI_DataType_F64 = 12,
  XLA_FFI_DataType_BF16 = 16
} XLA_FFI_DataType;

typedef struct XLA_FFI_Buffer {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_DataType dtype;
  void* data;
  int64_t rank;
  int64_t* dims;
} XLA_FFI_Buffer;

typedef enum { XLA_FFI_ArgType_BUFFER = 1 } XLA_FFI_ArgType;
typedef enum { XLA_FFI_RetType_BUFFER = 1 } XLA_FFI_RetType;
typedef enum { XLA_FFI_AttrType_ARRAY = 1, XLA_FFI_AttrType_DICTIONARY = 2, XLA_FFI_AttrType_SCALAR = 3, XLA_FFI_AttrType_STRING = 4 } XLA_FFI_AttrType;
typedef enum {
  XLA_FFI_ExecutionStage_INSTANTIATE = 0,
  XLA_FFI_ExecutionStage_PREPARE = 1,
  XLA_FFI_ExecutionStage_INITIALIZE = 2,
  XLA_FFI_ExecutionStage_EXECUTE = 3
} XLA_FFI_ExecutionStage;

typedef struct XLA_FFI_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_ArgType* types;
  void** args;
} XLA_FFI_Args;

typedef struct XLA_FFI_Rets {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_RetType* types;
  void** rets;
} XLA_FFI_Rets;

typedef struct XLA_FFI_ByteSpan {
  const char* ptr;
  size_t len;
} XLA_FFI_ByteSpan;

typedef struct XLA_FFI_Scalar {
  XLA_FFI_DataType dtype;
  void* value;
} XLA_FFI_Scalar;

typedef struct XLA_FFI_Attrs {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  int64_t size;
  XLA_FFI_AttrType* types;
  XLA_FFI_ByteSpan** names;
  void** attrs;
} XLA_FFI_Attrs;

typedef struct XLA_FFI_CallFrame {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  const XLA_FFI_Api* api;
  XLA_FFI_ExecutionContext* ctx;
  XLA_FFI_ExecutionStage stage;
  XLA_FFI_Args args;
  XLA_FFI_Rets rets;
  XLA_FFI_Attrs attrs;
  XLA_FFI_Future* future;
} XLA_FFI_CallFrame;

typedef XLA_FFI_Error* XLA_FFI_Handler(XLA_FFI_CallFrame* call_frame);

typedef struct XLA_FFI_Stream_Get_Args {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_ExecutionContext* ctx;
  void* stream; /* out */
} XLA_FFI_Stream_Get_Args;
typedef XLA_FFI_Error* XLA_FFI_Stream_Get(XLA_FFI_Stream_Get_Args* args);

/* only the leading members the shim calls through are declared; the real struct continues */
struct XLA_FFI_Api {
  size_t struct_size;
  XLA_FFI_Extension_Base* extension_start;
  XLA_FFI_Api_Version api_version;
  XLA_FFI_InternalApi* internal_api;
  XLA_FFI_Error_Create* XLA_FFI_Error_Create;
  void* XLA_FFI_Error_GetMessage;
  void* XLA_FFI_Error_Destroy;
  void* XLA_FFI_Handler_Register;
  XLA_FFI_Stream_Get* XLA_FFI_Stream_Get;
};

#ifdef __cplusplus
}
#endif
#endif /* XLA_FFI_API_C_API_H_ */
