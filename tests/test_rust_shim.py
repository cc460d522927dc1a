"""The Rust crate (rust/kogarashi-msm-b200) cannot be compiled here (no rustc / cargo), so its FFI surface is checked mechanically instead:
every prototype of include/kgr_msm.h must appear in the crate's `extern "C"` block with the same name, argument count, argument types
and return type, the #[repr(C)] job struct must match the C struct field by field, and every export the Python binding lists must be
declared in both."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_TO_RUST = {
    "int": "i32", "unsigned": "u32", "long": "c_long", "size_t": "usize", "uint64_t": "u64", "float": "f32", "double": "f64",
    "const char *": "*const c_char", "const int *": "*const i32", "const uint64_t *": "*const u64", "uint64_t *": "*mut u64", "const uint8_t *": "*const u8",
    "size_t *": "*mut usize", "float *": "*mut f32", "double *": "*mut f64", "uint32_t *": "*mut u32", "const void *": "*const c_void", "void *": "*mut c_void",
    "void **": "*mut *mut c_void",
    "kgr_bases_t *": "*mut kgr_bases_t", "const kgr_bases_t *": "*const kgr_bases_t", "kgr_bases_t **": "*mut *mut kgr_bases_t",
    "kgr_r1cs_t *": "*mut kgr_r1cs_t", "const kgr_r1cs_t *": "*const kgr_r1cs_t", "kgr_r1cs_t **": "*mut *mut kgr_r1cs_t",
    "kgr_vec_t *": "*mut kgr_vec_t", "const kgr_vec_t *": "*const kgr_vec_t", "kgr_vec_t **": "*mut *mut kgr_vec_t",
    "const kgr_msm_job_t *": "*const kgr_msm_job_t",
    "const uint32_t *const *": "*const *const u32", "const uint64_t *const *": "*const *const u64",
}


def _strip_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def _c_type(param):
    """'const uint64_t *xy' / 'float ms[9]' / 'const uint32_t *const row_ptr[3]' -> canonical C type (arrays decay to pointers)."""
    param = param.strip()
    array = re.search(r"\[\w*\]\s*$", param)
    if array:
        param = param[: array.start()].strip()
    m = re.match(r"^(.*?)(\w+)$", param)
    ctype = (m.group(1) if m and m.group(1).strip() else param).strip()
    if array:
        ctype += " *"
    ctype = re.sub(r"\s*\*\s*", " *", ctype)           # 'uint64_t*' -> 'uint64_t *'
    ctype = re.sub(r"\*\s+\*", "**", ctype)
    ctype = re.sub(r"\*\s*const\s*\*", "*const *", ctype)
    return re.sub(r"\s+", " ", ctype).strip()


def header_prototypes():
    text = _strip_comments(open(os.path.join(ROOT, "include", "kgr_msm.h")).read())
    protos = {}
    for m in re.finditer(r"(?m)^\s*(int|size_t|const char \*)\s*(kgr_\w+)\s*\(([^;{]*?)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = [] if args in ("void", "") else [_c_type(a) for a in args.split(",")]
        protos[name] = (ret.strip(), params)
    return protos


def rust_externs():
    text = open(os.path.join(ROOT, "rust", "kogarashi-msm-b200", "src", "lib.rs")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', text, flags=re.S).group(1)
    fns = {}
    for m in re.finditer(r"fn\s+(kgr_\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        name, args, ret = m.group(1), m.group(2), (m.group(3) or "()").strip()
        params = [re.sub(r"\s+", " ", a.split(":", 1)[1].strip()) for a in re.split(r",(?![^<]*>)", args) if a.strip()]
        fns[name] = (ret, params)
    return fns, text


def test_extern_block_matches_the_header():
    protos, (fns, _) = header_prototypes(), rust_externs()
    assert len(protos) >= 48, sorted(protos)
    assert set(fns) == set(protos), (sorted(set(protos) - set(fns)), sorted(set(fns) - set(protos)))
    for name, (ret, params) in protos.items():
        rret, rparams = fns[name]
        want_ret = {"int": "i32", "size_t": "usize", "const char *": "*const c_char"}[ret]
        assert rret == want_ret, (name, rret, want_ret)
        assert len(rparams) == len(params), (name, params, rparams)
        for i, (c, r) in enumerate(zip(params, rparams)):
            assert c in C_TO_RUST, (name, i, c)
            assert C_TO_RUST[c] == r, (name, i, c, r)


def test_job_struct_layout_matches():
    text = _strip_comments(open(os.path.join(ROOT, "include", "kgr_msm.h")).read())
    body = re.search(r"typedef struct kgr_msm_job \{(.*?)\} kgr_msm_job_t;", text, flags=re.S).group(1)
    c_fields = [(_c_type(f), re.search(r"(\w+)\s*$", f.strip()).group(1)) for f in body.split(";") if f.strip()]
    _, rust = rust_externs()
    rbody = re.search(r"pub struct kgr_msm_job_t \{(.*?)\}", rust, flags=re.S).group(1)
    r_fields = [(m.group(2).strip(), m.group(1)) for m in re.finditer(r"pub (\w+):\s*([^,]+),", rbody)]
    assert "#[repr(C)]\npub struct kgr_msm_job_t" in rust
    assert [n for _, n in c_fields] == [n for _, n in r_fields]
    for (ct, name), (rt, _) in zip(c_fields, r_fields):
        assert C_TO_RUST[ct] == rt, (name, ct, rt)


def test_python_binding_lists_every_export():
    from kogarashi_b200 import _lib
    protos = header_prototypes()
    assert set(_lib.EXPORTS) == set(protos), (sorted(set(protos) - set(_lib.EXPORTS)), sorted(set(_lib.EXPORTS) - set(protos)))


def test_shim_covers_what_the_prover_and_nova_call():
    """SURVEY H2 / VERDICT r1: G2Affine goes through the same generic entry point, the CRS is registered once, the fused prover call, NTT,
    Nova's resident vectors and a reusable pinned arena are bound; the stale 'G2 is out of scope' note is gone."""
    _, rust = rust_externs()
    for needle in ("impl GpuCurve for bn_254::G2Affine", "pub fn msm_curve_addition<C: GpuCurve>(bases: &[C], coeffs: &[C::Scalar]) -> C::Extended",
                   "pub struct GpuProver", "kgr_groth16_msms(", "pub struct PinnedArena", "pub struct DeviceVec", "pub struct R1csShapeGpu",
                   "kgr_nova_cross_term_device(", "pub fn ntt(", "pub fn msm_batch("):
        assert needle in rust, needle
    assert "out of scope" not in rust
