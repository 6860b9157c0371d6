// refrun.cc -- runs the reference's OWN native ops (the four shared objects inside
// core/custom_op/tensorflow_nms_car_3d-0.1.0-cp36-cp36m-linux_x86_64.whl) in this container,
// without TensorFlow.  TEST INFRASTRUCTURE ONLY (same rules as oracle/roi3d_oracle.c): it exists to
// pin the oracle against the real reference and to time the real reference on host cores.
//
// How: the wheel's libraries need libtensorflow_framework.so.2 (TF 2.2, cp36, old COW std::string
// ABI) which does not exist here.  Their op kernels, however, touch TensorFlow only through ~25
// out-of-line functions (OpKernelContext::input / allocate_output, TensorShape::dim_size,
// GetNodeAttr ...; see `nm -u`).  This file
//   1. maps a library's PT_LOAD segments itself and applies its dynamic relocations, resolving the
//      TensorFlow symbols to the stand-ins below and everything else (libc, libstdc++) through
//      dlsym -- WITHOUT running .init_array, so the REGISTER_OP / REGISTER_KERNEL_BUILDER static
//      initialisers (which need the real op registry) never execute;
//   2. finds the kernel factory lambda and <Op>::Compute in the library's .symtab, builds the
//      kernel object with the reference's own constructor, and calls the reference's own Compute
//      with a stand-in OpKernelContext that serves tensors in TF 2.2's in-memory layout
//      (TensorShapeRep: dims as uint16[6] at +0, dtype at +0xd, ndims at +0xe, tag at +0xf,
//      num_elements at +0x10; Tensor::buf_ at +0x18; TensorBuffer::data_ptr_ at +0x10).
// Nothing of the reference is copied into the repository: the libraries are read from the wheel
// where it lies (or from oracle/_ref/, a git-ignored extraction that travels to the GPU box).
//
// Build: g++ -O1 -shared -fPIC -D_GLIBCXX_USE_CXX11_ABI=0 (the reference was built with the old
// std::string ABI; libstdc++ still ships it).
#include <dlfcn.h>
#include <elf.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <string>
#include <vector>

#define REFRUN_API extern "C" __attribute__((visibility("default")))
#define TFSYM(name) asm(name) __attribute__((visibility("default"), used))

namespace refrun_types {

// ------------------------------------------------------------------------------------------
// stand-in TensorFlow objects
// ------------------------------------------------------------------------------------------
struct FakeShape {                 // tensorflow::TensorShapeRep, 24 bytes
    uint8_t buf[16];
    int64_t num_elements;
};
struct FakeBuffer { void *vtable; int64_t ref; void *data; };     // tensorflow::TensorBuffer
struct FakeTensor { FakeShape shape; FakeBuffer *buf; };          // tensorflow::Tensor, 32 bytes
struct FakeState { int code; std::string msg; };                  // tensorflow::Status::State

enum { DT_FLOAT = 1, DT_INT32 = 3 };

void shape_set(FakeShape *s, const long long *dims, size_t n, int dtype) {
    memset(s, 0, sizeof(*s));
    int64_t ne = 1;
    bool small = n <= 6;
    for (size_t i = 0; i < n; ++i) { ne *= dims[i]; small = small && dims[i] >= 0 && dims[i] < 65535; }
    if (small) {
        uint16_t *d16 = reinterpret_cast<uint16_t *>(s->buf);
        for (size_t i = 0; i < n; ++i) d16[i] = (uint16_t)dims[i];
        s->buf[15] = 0;                                            // REP16
    } else if (n <= 3) {
        uint32_t *d32 = reinterpret_cast<uint32_t *>(s->buf);
        for (size_t i = 0; i < n; ++i) d32[i] = (uint32_t)dims[i];
        s->buf[15] = 1;                                            // REP32
    } else {
        fprintf(stderr, "refrun: shape too large for the inline representations\n");
        abort();
    }
    s->buf[13] = (uint8_t)dtype;
    s->buf[14] = (uint8_t)n;
    s->num_elements = ne;
}
long long shape_dim(const FakeShape *s, int d) {
    if (s->buf[15] == 0) return reinterpret_cast<const uint16_t *>(s->buf)[d];
    return reinterpret_cast<const uint32_t *>(s->buf)[d];
}

struct Call {                        // state of the op invocation in flight (one per thread)
    std::vector<FakeTensor *> inputs;
    struct Out { FakeTensor t; FakeBuffer b; void *data; std::vector<long long> dims; };
    std::vector<Out *> outputs;
    bool failed = false;
    std::string error;
    std::string attr_method = "trilinear";
    float attr_extrapolation = 0.f, attr_iou_threshold = 0.5f;
    void *forced_out = nullptr;      // if set, output 0 is written here (caller-owned, large enough)
};
extern thread_local Call g_call;
thread_local Call g_call;

struct StatusRet { void *state; ~StatusRet() {} };                 // non-trivial dtor => returned via sret
struct StringRet { std::string s; };

}  // namespace refrun_types
using namespace refrun_types;

// ------------------------------------------------------------------------------------------
// stand-ins for the TensorFlow functions the kernels call (Itanium names from `nm -u`)
// ------------------------------------------------------------------------------------------
StatusRet tf_GetNodeAttr_str(const void *, const char *name, size_t len, std::string *value)
    TFSYM("_ZN10tensorflow11GetNodeAttrERKNS_9AttrSliceEN4absl11string_viewEPSs");
StatusRet tf_GetNodeAttr_str(const void *, const char *name, size_t len, std::string *value) {
    if (std::string(name, len) == "method_name") *value = g_call.attr_method;
    else { fprintf(stderr, "refrun: unexpected string attr %.*s\n", (int)len, name); abort(); }
    return StatusRet{nullptr};
}
StatusRet tf_GetNodeAttr_f(const void *, const char *name, size_t len, float *value)
    TFSYM("_ZN10tensorflow11GetNodeAttrERKNS_9AttrSliceEN4absl11string_viewEPf");
StatusRet tf_GetNodeAttr_f(const void *, const char *name, size_t len, float *value) {
    const std::string n(name, len);
    if (n == "extrapolation_value") *value = g_call.attr_extrapolation;
    else if (n == "iou_threshold") *value = g_call.attr_iou_threshold;
    else { fprintf(stderr, "refrun: unexpected float attr %s\n", n.c_str()); abort(); }
    return StatusRet{nullptr};
}
void tf_AttrSlice_ctor(void *, const void *) TFSYM("_ZN10tensorflow9AttrSliceC1ERKNS_7NodeDefE");
void tf_AttrSlice_ctor(void *, const void *) {}
void tf_OpKernel_ctor(void *, void *) TFSYM("_ZN10tensorflow8OpKernelC2EPNS_20OpKernelConstructionE");
void tf_OpKernel_ctor(void *, void *) {}
void tf_CheckNotInComputeAsync(void *, const char *) TFSYM("_ZN10tensorflow22CheckNotInComputeAsyncEPNS_15OpKernelContextEPKc");
void tf_CheckNotInComputeAsync(void *, const char *) {}

const FakeTensor *tf_ctx_input(void *, int i) TFSYM("_ZN10tensorflow15OpKernelContext5inputEi");
const FakeTensor *tf_ctx_input(void *, int i) { return g_call.inputs.at(i); }

StatusRet tf_ctx_allocate_output(void *, int idx, const FakeShape *shape, FakeTensor **out)
    TFSYM("_ZN10tensorflow15OpKernelContext15allocate_outputEiRKNS_11TensorShapeEPPNS_6TensorE");
StatusRet tf_ctx_allocate_output(void *, int idx, const FakeShape *shape, FakeTensor **out) {
    Call::Out *o = new Call::Out();
    const int nd = shape->buf[14];
    for (int d = 0; d < nd; ++d) o->dims.push_back(shape_dim(shape, d));
    const size_t bytes = (size_t)(shape->num_elements > 0 ? shape->num_elements : 0) * 4 + 64;
    o->data = (idx == 0 && g_call.forced_out) ? g_call.forced_out : aligned_alloc(64, (bytes + 63) / 64 * 64);
    o->b = FakeBuffer{nullptr, 1, o->data};
    o->t.shape = *shape;
    o->t.buf = &o->b;
    if ((size_t)idx >= g_call.outputs.size()) g_call.outputs.resize(idx + 1, nullptr);
    g_call.outputs[idx] = o;
    *out = &o->t;
    return StatusRet{nullptr};
}

static void record_failure(const char *file, int line, void *const *status) {
    g_call.failed = true;
    const FakeState *st = status ? static_cast<const FakeState *>(*status) : nullptr;
    g_call.error = std::string(file ? file : "?") + ":" + std::to_string(line) + ": " + (st ? st->msg : "");
}
void tf_ctx_CtxFailure(void *, const char *f, int l, void *const *s) TFSYM("_ZN10tensorflow15OpKernelContext10CtxFailureEPKciRKNS_6StatusE");
void tf_ctx_CtxFailure(void *, const char *f, int l, void *const *s) { record_failure(f, l, s); }
void tf_ctx_CtxFailureW(void *, const char *f, int l, void *const *s) TFSYM("_ZN10tensorflow15OpKernelContext21CtxFailureWithWarningEPKciRKNS_6StatusE");
void tf_ctx_CtxFailureW(void *, const char *f, int l, void *const *s) { record_failure(f, l, s); }
void tf_con_CtxFailure(void *, const char *f, int l, void *const *s) TFSYM("_ZN10tensorflow20OpKernelConstruction10CtxFailureEPKciRKNS_6StatusE");
void tf_con_CtxFailure(void *, const char *f, int l, void *const *s) { record_failure(f, l, s); }
void tf_con_CtxFailureW(void *, const char *f, int l, void *const *s) TFSYM("_ZN10tensorflow20OpKernelConstruction21CtxFailureWithWarningEPKciRKNS_6StatusE");
void tf_con_CtxFailureW(void *, const char *f, int l, void *const *s) { record_failure(f, l, s); }
void tf_ctx_SetStatus(void *, void *const *s) TFSYM("_ZN10tensorflow15OpKernelContext9SetStatusERKNS_6StatusE");
void tf_ctx_SetStatus(void *, void *const *s) { if (s && *s) record_failure("SetStatus", 0, s); }

void tf_Status_ctor(void **self, int code, const char *msg, size_t len) TFSYM("_ZN10tensorflow6StatusC1ENS_5error4CodeEN4absl11string_viewE");
void tf_Status_ctor(void **self, int code, const char *msg, size_t len) { *self = new FakeState{code, std::string(msg, len)}; }

void tf_Shape_ctor(FakeShape *self, const long long *dims, size_t n) TFSYM("_ZN10tensorflow15TensorShapeBaseINS_11TensorShapeEEC1EN4absl4SpanIKxEE");
void tf_Shape_ctor(FakeShape *self, const long long *dims, size_t n) { shape_set(self, dims, n, 0); }
void tf_Shape_dtor_ool(FakeShape *) TFSYM("_ZN10tensorflow14TensorShapeRep19DestructorOutOfLineEv");
void tf_Shape_dtor_ool(FakeShape *) {}
void tf_Shape_slowcopy(FakeShape *self, const FakeShape *o) TFSYM("_ZN10tensorflow14TensorShapeRep12SlowCopyFromERKS0_");
void tf_Shape_slowcopy(FakeShape *self, const FakeShape *o) { *self = *o; }
void tf_Shape_CheckDimsEqual(const FakeShape *, int) TFSYM("_ZNK10tensorflow11TensorShape14CheckDimsEqualEi");
void tf_Shape_CheckDimsEqual(const FakeShape *, int) {}
void tf_Shape_CheckDimsAtLeast(const FakeShape *, int) TFSYM("_ZNK10tensorflow11TensorShape16CheckDimsAtLeastEi");
void tf_Shape_CheckDimsAtLeast(const FakeShape *, int) {}
long long tf_Shape_dim_size(const FakeShape *s, int d) TFSYM("_ZNK10tensorflow15TensorShapeBaseINS_11TensorShapeEE8dim_sizeEi");
long long tf_Shape_dim_size(const FakeShape *s, int d) { return shape_dim(s, d); }
StringRet tf_Shape_DebugString(const FakeShape *) TFSYM("_ZNK10tensorflow14TensorShapeRep11DebugStringEv");
StringRet tf_Shape_DebugString(const FakeShape *) { return StringRet{"[shape]"}; }

void tf_Tensor_CheckTypeAligned(const FakeTensor *, int) TFSYM("_ZNK10tensorflow6Tensor21CheckTypeAndIsAlignedENS_8DataTypeE");
void tf_Tensor_CheckTypeAligned(const FakeTensor *, int) {}
void tf_Tensor_CheckType(const FakeTensor *, int) TFSYM("_ZNK10tensorflow6Tensor9CheckTypeENS_8DataTypeE");
void tf_Tensor_CheckType(const FakeTensor *, int) {}
void tf_Tensor_CheckSingle(const FakeTensor *) TFSYM("_ZNK10tensorflow6Tensor30CheckIsAlignedAndSingleElementEv");
void tf_Tensor_CheckSingle(const FakeTensor *) {}

StringRet tf_StrCat1(const void *) TFSYM("_ZN10tensorflow7strings6StrCatERKNS0_8AlphaNumE");
StringRet tf_StrCat1(const void *) { return StringRet{"<refrun: op validation failed>"}; }
StringRet tf_StrCat2(const void *, const void *) TFSYM("_ZN10tensorflow7strings6StrCatERKNS0_8AlphaNumES3_");
StringRet tf_StrCat2(const void *, const void *) { return StringRet{"<refrun: op validation failed>"}; }
StringRet tf_CatPieces(const void *, size_t) TFSYM("_ZN10tensorflow7strings8internal9CatPiecesESt16initializer_listIN4absl11string_viewEE");
StringRet tf_CatPieces(const void *, size_t) { return StringRet{"<refrun: op validation failed>"}; }
StringRet tf_DataTypeString(int) TFSYM("_ZN10tensorflow14DataTypeStringENS_8DataTypeE");
StringRet tf_DataTypeString(int) { return StringRet{"dtype"}; }
char *tf_FastInt32(int v, char *buf) TFSYM("_ZN10tensorflow7strings21FastInt32ToBufferLeftEiPc");
char *tf_FastInt32(int v, char *buf) { return buf + sprintf(buf, "%d", v); }

// data symbols referenced by relocations (never dereferenced on the paths we run)
extern "C" {
__attribute__((visibility("default"))) const char *refrun_DEVICE_CPU asm("_ZN10tensorflow10DEVICE_CPUE") = "CPU";
__attribute__((visibility("default"))) const char *refrun_DEVICE_GPU asm("_ZN10tensorflow10DEVICE_GPUE") = "GPU";
__attribute__((visibility("default"))) void *refrun_ti_OpKernel[4] asm("_ZTIN10tensorflow8OpKernelE") = {nullptr, nullptr, nullptr, nullptr};
__attribute__((visibility("default"))) void *refrun_vt_PtrFactory[8]
    asm("_ZTVN10tensorflow14kernel_factory17OpKernelRegistrar18PtrOpKernelFactoryE") = {nullptr};
}

namespace {

// every other TensorFlow / CUDA symbol: must never be reached
const char *g_trap_names[512];
int g_ntraps = 0;
extern "C" void refrun_trap() {
    fprintf(stderr, "refrun: the reference called a TensorFlow/CUDA function that has no stand-in\n");
    abort();
}

// ------------------------------------------------------------------------------------------
// minimal ELF64 loader (no .init_array, no TLS, x86-64 relocation types used by these files)
// ------------------------------------------------------------------------------------------
struct Lib {
    uint8_t *base = nullptr;
    size_t span = 0;
    std::vector<uint8_t> file;
    const Elf64_Sym *symtab = nullptr;
    size_t nsyms = 0;
    const char *strtab = nullptr;

    void *find(const char *must1, const char *must2 = nullptr, int nth = 0) const {
        int seen = 0;
        for (size_t i = 0; i < nsyms; ++i) {
            const Elf64_Sym &s = symtab[i];
            if (ELF64_ST_TYPE(s.st_info) != STT_FUNC || s.st_shndx == SHN_UNDEF) continue;
            const char *n = strtab + s.st_name;
            if (strstr(n, must1) && (!must2 || strstr(n, must2)) && seen++ == nth) return base + s.st_value;
        }
        return nullptr;
    }
    void *find_object(const char *name) const {
        for (size_t i = 0; i < nsyms; ++i) {
            const Elf64_Sym &s = symtab[i];
            if (s.st_shndx == SHN_UNDEF) continue;
            if (!strcmp(strtab + s.st_name, name)) return base + s.st_value;
        }
        return nullptr;
    }
};

bool load_lib(const char *path, Lib *L, std::string *err) {
    FILE *f = fopen(path, "rb");
    if (!f) { *err = std::string("cannot open ") + path; return false; }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    L->file.resize(sz);
    if (fread(L->file.data(), 1, sz, f) != (size_t)sz) { fclose(f); *err = "short read"; return false; }
    fclose(f);
    const uint8_t *d = L->file.data();
    const Elf64_Ehdr *eh = reinterpret_cast<const Elf64_Ehdr *>(d);
    if (memcmp(eh->e_ident, ELFMAG, SELFMAG) || eh->e_machine != EM_X86_64) { *err = "not an x86-64 ELF"; return false; }
    const Elf64_Phdr *ph = reinterpret_cast<const Elf64_Phdr *>(d + eh->e_phoff);
    uint64_t hi = 0;
    const Elf64_Dyn *dyn = nullptr;
    for (int i = 0; i < eh->e_phnum; ++i) {
        if (ph[i].p_type == PT_LOAD && ph[i].p_vaddr + ph[i].p_memsz > hi) hi = ph[i].p_vaddr + ph[i].p_memsz;
        if (ph[i].p_type == PT_TLS) { *err = "TLS segment not supported"; return false; }
    }
    L->span = (hi + 4095) & ~4095ull;
    L->base = static_cast<uint8_t *>(mmap(nullptr, L->span, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    if (L->base == MAP_FAILED) { *err = "mmap failed"; return false; }
    for (int i = 0; i < eh->e_phnum; ++i) {
        if (ph[i].p_type == PT_LOAD) memcpy(L->base + ph[i].p_vaddr, d + ph[i].p_offset, ph[i].p_filesz);
        if (ph[i].p_type == PT_DYNAMIC) dyn = reinterpret_cast<const Elf64_Dyn *>(L->base + ph[i].p_vaddr);
    }
    if (!dyn) { *err = "no PT_DYNAMIC"; return false; }
    const Elf64_Sym *dsym = nullptr;
    const char *dstr = nullptr;
    const Elf64_Rela *rela = nullptr, *jmprel = nullptr;
    size_t relasz = 0, jmpsz = 0;
    for (const Elf64_Dyn *p = dyn; p->d_tag != DT_NULL; ++p) {
        switch (p->d_tag) {
        case DT_SYMTAB: dsym = reinterpret_cast<const Elf64_Sym *>(L->base + p->d_un.d_ptr); break;
        case DT_STRTAB: dstr = reinterpret_cast<const char *>(L->base + p->d_un.d_ptr); break;
        case DT_RELA: rela = reinterpret_cast<const Elf64_Rela *>(L->base + p->d_un.d_ptr); break;
        case DT_RELASZ: relasz = p->d_un.d_val; break;
        case DT_JMPREL: jmprel = reinterpret_cast<const Elf64_Rela *>(L->base + p->d_un.d_ptr); break;
        case DT_PLTRELSZ: jmpsz = p->d_un.d_val; break;
        default: break;
        }
    }
    auto resolve = [&](const Elf64_Sym &s) -> uint64_t {
        if (s.st_shndx != SHN_UNDEF) return reinterpret_cast<uint64_t>(L->base + s.st_value);
        const char *name = dstr + s.st_name;
        if (void *p = dlsym(RTLD_DEFAULT, name)) return reinterpret_cast<uint64_t>(p);
        if (ELF64_ST_BIND(s.st_info) == STB_WEAK) return 0;
        if (g_ntraps < 512) g_trap_names[g_ntraps++] = name;
        if (getenv("REFRUN_DEBUG")) fprintf(stderr, "refrun: unresolved -> trap: %s\n", name);
        return reinterpret_cast<uint64_t>(&refrun_trap);
    };
    auto apply = [&](const Elf64_Rela *r, size_t bytes) -> bool {
        for (size_t i = 0; i < bytes / sizeof(Elf64_Rela); ++i) {
            uint64_t *where = reinterpret_cast<uint64_t *>(L->base + r[i].r_offset);
            const Elf64_Sym &s = dsym[ELF64_R_SYM(r[i].r_info)];
            switch (ELF64_R_TYPE(r[i].r_info)) {
            case R_X86_64_RELATIVE: *where = reinterpret_cast<uint64_t>(L->base) + r[i].r_addend; break;
            case R_X86_64_GLOB_DAT:
            case R_X86_64_JUMP_SLOT: *where = resolve(s); break;
            case R_X86_64_64: *where = resolve(s) + r[i].r_addend; break;
            default: *err = "unsupported relocation type " + std::to_string(ELF64_R_TYPE(r[i].r_info)); return false;
            }
        }
        return true;
    };
    if (rela && !apply(rela, relasz)) return false;
    if (jmprel && !apply(jmprel, jmpsz)) return false;
    for (int i = 0; i < eh->e_phnum; ++i) {
        if (ph[i].p_type == PT_LOAD && (ph[i].p_flags & PF_X)) {
            const uint64_t a = ph[i].p_vaddr & ~4095ull, b = (ph[i].p_vaddr + ph[i].p_memsz + 4095) & ~4095ull;
            mprotect(L->base + a, b - a, PROT_READ | PROT_EXEC);
        }
    }
    // .symtab / .strtab from the section headers (the files are not stripped)
    const Elf64_Shdr *sh = reinterpret_cast<const Elf64_Shdr *>(d + eh->e_shoff);
    for (int i = 0; i < eh->e_shnum; ++i) {
        if (sh[i].sh_type == SHT_SYMTAB) {
            L->symtab = reinterpret_cast<const Elf64_Sym *>(d + sh[i].sh_offset);
            L->nsyms = sh[i].sh_size / sizeof(Elf64_Sym);
            L->strtab = reinterpret_cast<const char *>(d + sh[sh[i].sh_link].sh_offset);
        }
    }
    if (!L->symtab) { *err = "no .symtab"; return false; }
    return true;
}

typedef void *(*factory_fn)(void *construction);
typedef void (*compute_fn)(void *self, void *ctx);

struct Op {
    Lib lib;
    factory_fn factory = nullptr;
    compute_fn compute = nullptr;
    bool ok = false;
};
Op g_car, g_gi, g_gb, g_nms;
thread_local std::string g_error;

bool open_op(Op *op, const char *path, const char *compute_must) {
    if (!load_lib(path, &op->lib, &g_error)) return false;
    op->compute = reinterpret_cast<compute_fn>(op->lib.find(compute_must, "7ComputeEPN"));
    if (!op->compute) { g_error = std::string("Compute not found in ") + path; return false; }
    // the kernel factory: `[](OpKernelConstruction* c) -> OpKernel* { return new Op(c); }`
    for (int nth = 0;; ++nth) {
        void *f = op->lib.find("OpKernelConstructionEE_4_FUN", nullptr, nth);
        if (!f) break;
        op->factory = reinterpret_cast<factory_fn>(f);
        // several kernels may live in one library (NMS.so also holds the dead 2-D GPU op): pick the
        // factory whose object dispatches to the Compute we want
        alignas(64) static uint8_t construction[4096];
        memset(construction, 0, sizeof(construction));
        g_call.failed = false;
        void *obj = op->factory(construction);
        if (!obj) continue;
        void **vtable = *reinterpret_cast<void ***>(obj);
        bool match = false;
        for (int s = 0; s < 8 && !match; ++s) match = vtable[s] == reinterpret_cast<void *>(op->compute);
        if (match) { op->ok = true; return true; }
    }
    g_error = std::string("kernel factory not found in ") + path;
    return false;
}

FakeTensor *make_tensor(std::vector<FakeBuffer *> &bufs, const void *data, std::initializer_list<long long> dims, int dtype) {
    FakeTensor *t = new FakeTensor();
    std::vector<long long> d(dims);
    shape_set(&t->shape, d.data(), d.size(), dtype);
    FakeBuffer *b = new FakeBuffer{nullptr, 1, const_cast<void *>(data)};
    bufs.push_back(b);
    t->buf = b;
    return t;
}

// run one op: fresh kernel object (reference constructor), reference Compute, collect output 0
int run(Op *op, std::initializer_list<FakeTensor *> inputs, void *forced_out, void **out_data, std::vector<long long> *out_dims) {
    if (!op->ok) { g_error = "library not loaded"; return -1; }
    alignas(64) uint8_t construction[4096], context[8192];      // stand-ins for OpKernelConstruction / OpKernelContext
    memset(construction, 0, sizeof(construction));
    memset(context, 0, sizeof(context));
    g_call.inputs.assign(inputs.begin(), inputs.end());
    g_call.outputs.clear();
    g_call.failed = false;
    g_call.forced_out = forced_out;
    void *kernel = op->factory(construction);
    if (g_call.failed || !kernel) { g_error = "kernel construction failed: " + g_call.error; return -2; }
    op->compute(kernel, context);
    for (FakeTensor *t : g_call.inputs) { delete t; }
    g_call.inputs.clear();
    if (g_call.failed) { g_error = "Compute failed: " + g_call.error; return -3; }
    if (g_call.outputs.empty() || !g_call.outputs[0]) { g_error = "no output allocated"; return -4; }
    if (out_data) *out_data = g_call.outputs[0]->data;
    if (out_dims) *out_dims = g_call.outputs[0]->dims;
    return 0;
}

void free_outputs(bool keep_data0) {
    for (size_t i = 0; i < g_call.outputs.size(); ++i) {
        Call::Out *o = g_call.outputs[i];
        if (!o) continue;
        if (!(i == 0 && (keep_data0 || g_call.forced_out))) free(o->data);
        delete o;
    }
    g_call.outputs.clear();
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C API (ctypes)
// ------------------------------------------------------------------------------------------
REFRUN_API const char *refrun_last_error() { return g_error.c_str(); }

REFRUN_API int refrun_open(const char *car, const char *gi, const char *gb, const char *nms) {
    if (!g_car.ok && !open_op(&g_car, car, "17CropAndResize3DOp")) return -1;
    if (!g_gi.ok && !open_op(&g_gi, gi, "26CropAndResize3DGradImageOp")) return -2;
    if (!g_gb.ok && !open_op(&g_gb, gb, "26CropAndResize3DGradBoxesOp")) return -3;
    if (!g_nms.ok && !open_op(&g_nms, nms, "21NonMaxSuppression3DOp")) return -4;
    return 0;
}

REFRUN_API int refrun_car3d_fwd(const float *image, int B, int H, int W, int D, int C, const float *boxes,
                                const int *box_index, int n, int ph, int pw, int pd, int method, float ext, float *crops) {
    std::vector<FakeBuffer *> bufs;
    const int crop[3] = {ph, pw, pd};
    g_call.attr_method = method ? "nearest" : "trilinear";
    g_call.attr_extrapolation = ext;
    const int rc = run(&g_car, {make_tensor(bufs, image, {B, H, W, D, C}, DT_FLOAT), make_tensor(bufs, boxes, {n, 6}, DT_FLOAT),
                                make_tensor(bufs, box_index, {n}, DT_INT32), make_tensor(bufs, crop, {3}, DT_INT32)},
                       crops, nullptr, nullptr);
    free_outputs(false);
    for (FakeBuffer *b : bufs) delete b;
    return rc;
}

REFRUN_API int refrun_car3d_grad_image(const float *grads, const float *boxes, const int *box_ind, int n, int ph, int pw, int pd,
                                       int B, int H, int W, int D, int C, int method, float *out) {
    std::vector<FakeBuffer *> bufs;
    const int size[5] = {B, H, W, D, C};
    g_call.attr_method = method ? "nearest" : "trilinear";
    const int rc = run(&g_gi, {make_tensor(bufs, grads, {n, ph, pw, pd, C}, DT_FLOAT), make_tensor(bufs, boxes, {n, 6}, DT_FLOAT),
                               make_tensor(bufs, box_ind, {n}, DT_INT32), make_tensor(bufs, size, {5}, DT_INT32)},
                       out, nullptr, nullptr);
    free_outputs(false);
    for (FakeBuffer *b : bufs) delete b;
    return rc;
}

REFRUN_API int refrun_car3d_grad_boxes(const float *grads, const float *image, int B, int H, int W, int D, int C,
                                       const float *boxes, const int *box_ind, int n, int ph, int pw, int pd, float *out) {
    std::vector<FakeBuffer *> bufs;
    g_call.attr_method = "trilinear";
    const int rc = run(&g_gb, {make_tensor(bufs, grads, {n, ph, pw, pd, C}, DT_FLOAT), make_tensor(bufs, image, {B, H, W, D, C}, DT_FLOAT),
                               make_tensor(bufs, boxes, {n, 6}, DT_FLOAT), make_tensor(bufs, box_ind, {n}, DT_INT32)},
                       out, nullptr, nullptr);
    free_outputs(false);
    for (FakeBuffer *b : bufs) delete b;
    return rc;
}

// the reference's own IOU<float>(TTypes<float,2>::ConstTensor boxes, int i, int j) (NMS.so@0xb500): the
// Eigen TensorMap {data, dim0, dim1} is a 24-byte aggregate passed in memory
struct IouTensorMap { const float *data; long d0, d1; };
REFRUN_API int refrun_iou_pairs(const float *boxes, int n, const int *ii, const int *jj, int npairs, float *out) {
    if (!g_nms.ok) { g_error = "library not loaded"; return -1; }
    typedef float (*iou_fn)(IouTensorMap, int, int);
    static iou_fn fn = nullptr;
    if (!fn) fn = reinterpret_cast<iou_fn>(g_nms.lib.find("IOUIf"));
    if (!fn) { g_error = "IOU<float> not found"; return -2; }
    const IouTensorMap m{boxes, n, 6};
    for (int p = 0; p < npairs; ++p) out[p] = fn(m, ii[p], jj[p]);
    return 0;
}

// returns the number of selected indices (>= 0) or a negative error
REFRUN_API int refrun_nms3d(const float *boxes, const float *scores, int n, int max_out, float thr, int *selected, int capacity) {
    std::vector<FakeBuffer *> bufs;
    g_call.attr_iou_threshold = thr;
    void *data = nullptr;
    std::vector<long long> dims;
    const int rc = run(&g_nms, {make_tensor(bufs, boxes, {n, 6}, DT_FLOAT), make_tensor(bufs, scores, {n}, DT_FLOAT),
                                make_tensor(bufs, &max_out, {}, DT_INT32)},
                       nullptr, &data, &dims);
    int m = rc;
    if (rc == 0) {
        m = dims.empty() ? 0 : (int)dims[0];
        if (m > capacity) { g_error = "selected buffer too small"; m = -5; }
        else memcpy(selected, data, sizeof(int) * (size_t)m);
    }
    free_outputs(false);
    for (FakeBuffer *b : bufs) delete b;
    return m;
}
