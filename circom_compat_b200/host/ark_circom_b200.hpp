// ark_circom_b200.hpp - C++ host-side mirror of the ark-circom proving interface, on top of the C ABI (include/b2groth.h).
//
// The reference is compiled Rust and no Rust toolchain exists in the build image, so the host layer a user links
// against is C++ with the reference's names, argument meaning and error behaviour:
//
//   ark_circom::read_zkey(reader)                      <- /root/reference/src/zkey.rs:53-60 (+ BinFile :73-133, :151-196)
//   ark_circom::ProvingKey / ConstraintMatrices        <- ProvingKey<Bn254> (zkey.rs:121-130) / ConstraintMatrices<Fr> (zkey.rs:181-193)
//   ark_circom::CircomReduction::witness_map_from_matrices      <- src/circom/qap.rs:23-88
//   ark_circom::Groth16::create_proof_with_reduction_and_matrices <- call sites src/zkey.rs:903-912, benches/groth16.rs:52-61
//   ark_circom::Groth16::prove                         <- src/zkey.rs:866 (draws r then s)
//   ark_circom::Groth16::process_vk / verify_with_processed_vk / verify <- src/zkey.rs:868-870, tests/groth16.rs:33-35
//                                                         (host pairing, ark_circom_verifier.hpp; no GPU involved)
//   ark_circom::read_wtns                              <- snarkjs .wtns (test-vectors/circuit2_js/witness.wtns; the reference
//                                                         computes witnesses with WASM instead, out of scope here)
// Parsing and key handling stay on the host; every field/curve operation of the proof runs in libb2groth.so.
// Header-only; link with -lb2groth.
#pragma once
#include <cstdint>
#include <cstring>
#include <istream>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b2groth.h"

namespace ark_circom {

// ---------------------------------------------------------------------------------------------- errors
struct SerializationError : std::runtime_error { using std::runtime_error::runtime_error; };                 // zkey.rs:43
struct SynthesisError : std::runtime_error { using std::runtime_error::runtime_error; };
struct PolynomialDegreeTooLarge : SynthesisError { PolynomialDegreeTooLarge() : SynthesisError("PolynomialDegreeTooLarge") {} };  // qap.rs:31
struct DeviceError : std::runtime_error { using std::runtime_error::runtime_error; };

inline void check(int rc) {
    if (rc == B2G_OK) return;
    if (rc == B2G_E_DOMAIN) throw PolynomialDegreeTooLarge();
    throw DeviceError(std::string("b2groth error ") + std::to_string(rc) + ": " + b2g_last_error());
}

// ---------------------------------------------------------------------------------------------- Fr (host side: conversions only)
typedef unsigned __int128 u128;
struct BigInt256 { uint64_t l[4]; };

namespace detail {
static const uint64_t FR_P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FQ_P[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FR_INV = 0xc2e1f593efffffffULL;
static const uint64_t FR_R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};

inline bool geq(const uint64_t a[4], const uint64_t p[4]) {
    for (int i = 3; i >= 0; i--) { if (a[i] > p[i]) return true; if (a[i] < p[i]) return false; }
    return true;
}
// Montgomery product mod r (host, used for encodings only - never for the proof)
inline void fr_mont_mul(uint64_t out[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_INV;
        c = (u128)m * FR_P[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * FR_P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; c >>= 64;
        t[4] = t[5] + (uint64_t)c;
    }
    if (t[4] || geq(t, FR_P)) { u128 br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - FR_P[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; } }
    memcpy(out, t, 32);
}
}  // namespace detail

// Fr as arkworks keeps it: 4 x u64 Montgomery limbs (Fp256<MontBackend>)
struct Fr {
    uint64_t l[4] = {0, 0, 0, 0};
    static Fr new_unchecked(const BigInt256& b) { Fr f; memcpy(f.l, b.l, 32); return f; }          // limbs ARE the residue
    static Fr from_bigint(const BigInt256& b) {                                                     // canonical -> Montgomery
        if (detail::geq(b.l, detail::FR_P)) throw std::invalid_argument("Fr::from_bigint: not reduced");
        Fr f; detail::fr_mont_mul(f.l, b.l, detail::FR_R2); return f;
    }
    static Fr from_u64(uint64_t v) { BigInt256 b = {{v, 0, 0, 0}}; return from_bigint(b); }
    BigInt256 into_bigint() const { const uint64_t one[4] = {1, 0, 0, 0}; BigInt256 b; detail::fr_mont_mul(b.l, l, one); return b; }
    bool is_zero() const { return !(l[0] | l[1] | l[2] | l[3]); }
    bool operator==(const Fr& o) const { return !memcmp(l, o.l, 32); }
    // Fr::rand of ark-ff 0.5: 4 limbs from the rng, top two bits cleared, rejected if >= r,
    // interpreted as the Montgomery residue
    template <class Rng> static Fr rand(Rng& rng) {
        for (;;) {
            BigInt256 b;
            for (int i = 0; i < 4; i++) b.l[i] = rng();
            b.l[3] &= 0x3fffffffffffffffULL;
            if (!detail::geq(b.l, detail::FR_P)) return new_unchecked(b);
        }
    }
};

struct G1Affine { uint64_t x[4], y[4]; bool is_infinity() const { uint64_t o = 0; for (int i = 0; i < 4; i++) o |= x[i] | y[i]; return !o; } };   // Montgomery, zeros = infinity
struct G2Affine { uint64_t x0[4], x1[4], y0[4], y1[4]; };
static_assert(sizeof(G1Affine) == 64 && sizeof(G2Affine) == 128, "zkey point layout");

struct VerifyingKey { G1Affine alpha_g1; G2Affine beta_g2, gamma_g2, delta_g2; std::vector<G1Affine> gamma_abc_g1; };

// Device copies (fixed-base tables, CSR matrices) belong to the host object they were made from: the slot is a member
// of ProvingKey / ConstraintMatrices, so the gigabytes of HBM are freed when that object dies, and a copy or an
// assignment - the ways a *different* key can come to live at the same address - start with an empty slot.  The device
// copy is a snapshot taken at first use: after mutating a key in place call release_device().
class DeviceSlot {
public:
    DeviceSlot() = default;
    DeviceSlot(const DeviceSlot&) {}                                   // a copy has no device state of its own yet
    DeviceSlot(DeviceSlot&& o) noexcept : handles_(std::move(o.handles_)), free_(o.free_) { o.handles_.clear(); }
    DeviceSlot& operator=(const DeviceSlot&) { release(); return *this; }   // new contents => stale tables must go
    DeviceSlot& operator=(DeviceSlot&& o) noexcept { if (this != &o) { release(); handles_ = std::move(o.handles_); free_ = o.free_; o.handles_.clear(); } return *this; }
    ~DeviceSlot() { release(); }
    void release() const { for (auto& kv : handles_) if (kv.second && free_) free_(kv.second); handles_.clear(); }
    void* find(const void* ctx, uint32_t tag) const { auto it = handles_.find({ctx, tag}); return it == handles_.end() ? nullptr : it->second; }
    void put(const void* ctx, uint32_t tag, void* h, void (*free_fn)(void*)) const { handles_[{ctx, tag}] = h; free_ = free_fn; }
private:
    mutable std::map<std::pair<const void*, uint32_t>, void*> handles_;       // (b2g_ctx, variant) -> b2g_pk* / b2g_mat*
    mutable void (*free_)(void*) = nullptr;
};

struct ProvingKey {                                     // ProvingKey<Bn254>, src/zkey.rs:121-130
    VerifyingKey vk;
    G1Affine beta_g1, delta_g1;
    std::vector<G1Affine> a_query, b_g1_query, h_query, l_query;
    std::vector<G2Affine> b_g2_query;
    DeviceSlot device;                                  // see DeviceSlot
    void release_device() const { device.release(); }
};

typedef std::vector<std::vector<std::pair<Fr, size_t>>> Matrix;     // rows of (coeff, index): src/zkey.rs:168

struct ConstraintMatrices {                             // src/zkey.rs:181-193
    size_t num_instance_variables = 0, num_witness_variables = 0, num_constraints = 0;
    size_t a_num_non_zero = 0, b_num_non_zero = 0, c_num_non_zero = 0;
    Matrix a, b, c;
    DeviceSlot device;
    void release_device() const { device.release(); }
};

struct Proof {                                          // Proof<Bn254>; coordinates canonical little-endian
    uint8_t bytes[256];                                 // A.x A.y B.x.c0 B.x.c1 B.y.c0 B.y.c1 C.x C.y
    std::string hex() const { static const char* d = "0123456789abcdef"; std::string s; for (uint8_t b : bytes) { s += d[b >> 4]; s += d[b & 15]; } return s; }
};

// ---------------------------------------------------------------------------------------------- zkey reader (host)
namespace detail {
struct Section { uint64_t position, size; };
inline void read_exact(std::istream& r, void* dst, size_t n) {
    r.read(reinterpret_cast<char*>(dst), (std::streamsize)n);
    if ((size_t)r.gcount() != n) throw SerializationError("unexpected end of zkey");
}
template <class T> inline T read_le(std::istream& r) { T v; read_exact(r, &v, sizeof(T)); return v; }   // x86: little-endian host
}  // namespace detail

class BinFile {                                         // src/zkey.rs:62-101
public:
    explicit BinFile(std::istream& reader) : r_(reader) {
        char magic[4]; detail::read_exact(r_, magic, 4);
        ftype_.assign(magic, 4);
        version_ = detail::read_le<uint32_t>(r_);
        uint32_t nsec = detail::read_le<uint32_t>(r_);
        for (uint32_t i = 0; i < nsec; i++) {
            uint32_t id = detail::read_le<uint32_t>(r_);
            uint64_t len = detail::read_le<uint64_t>(r_);
            sections_[id].push_back({(uint64_t)r_.tellg(), len});
            r_.seekg((std::streamoff)len, std::ios::cur);
            if (!r_) throw SerializationError("truncated zkey section table");
        }
        if (ftype_ != "zkey") throw SerializationError("not a zkey file");
    }

    ProvingKey proving_key() {                          // src/zkey.rs:103-133
        Header h = groth_header();
        ProvingKey pk;
        pk.vk.alpha_g1 = h.alpha_g1; pk.vk.beta_g2 = h.beta_g2; pk.vk.gamma_g2 = h.gamma_g2; pk.vk.delta_g2 = h.delta_g2;
        pk.beta_g1 = h.beta_g1; pk.delta_g1 = h.delta_g1;
        pk.vk.gamma_abc_g1 = g1_section(h.n_public + 1, 3);
        pk.a_query = g1_section(h.n_vars, 5);
        pk.b_g1_query = g1_section(h.n_vars, 6);
        pk.b_g2_query = g2_section(h.n_vars, 7);
        pk.l_query = g1_section(h.n_vars - h.n_public - 1, 8);
        pk.h_query = g1_section(h.domain_size, 9);
        return pk;
    }

    ConstraintMatrices matrices() {                     // src/zkey.rs:151-196
        Header h = groth_header();
        seek(4);
        uint32_t ncoef = detail::read_le<uint32_t>(r_);
        std::vector<Matrix> m(2, Matrix(h.domain_size));
        uint32_t max_c = 0;
        const uint64_t one[4] = {1, 0, 0, 0};
        for (uint32_t i = 0; i < ncoef; i++) {
            uint32_t matrix = detail::read_le<uint32_t>(r_), constraint = detail::read_le<uint32_t>(r_), signal = detail::read_le<uint32_t>(r_);
            BigInt256 raw; detail::read_exact(r_, raw.l, 32);
            if (matrix > 1 || constraint >= h.domain_size) throw SerializationError("bad coefficient record");
            // stored = v * R^2; the reader strips one R (zkey.rs:320-325): Montgomery residue of v = stored * R^-1
            Fr v; detail::fr_mont_mul(v.l, raw.l, one);
            if (constraint > max_c) max_c = constraint;
            m[matrix][constraint].push_back({v, (size_t)signal});
        }
        if (max_c < h.n_public) throw SerializationError("malformed zkey: no constraints");
        size_t nc = max_c - h.n_public;                 // zkey.rs:171
        for (auto& mm : m) mm.resize(nc);               // public-input rows dropped, arkworks re-adds them (qap.rs:46-50)
        ConstraintMatrices cm;
        cm.num_instance_variables = h.n_public + 1; cm.num_witness_variables = h.n_vars - h.n_public - 1; cm.num_constraints = nc;
        cm.a = std::move(m[0]); cm.b = std::move(m[1]);
        for (auto& row : cm.a) cm.a_num_non_zero += row.size();
        for (auto& row : cm.b) cm.b_num_non_zero += row.size();
        return cm;
    }

    struct Header { uint32_t n_vars, n_public, domain_size; G1Affine alpha_g1, beta_g1, delta_g1; G2Affine beta_g2, gamma_g2, delta_g2; };
    Header groth_header() {                             // src/zkey.rs:282-318
        seek(2);
        Header h;
        uint32_t n8q = detail::read_le<uint32_t>(r_);
        if (n8q != 32) throw SerializationError("unsupported base field size");
        uint64_t q[4]; detail::read_exact(r_, q, 32);
        uint32_t n8r = detail::read_le<uint32_t>(r_);
        if (n8r != 32) throw SerializationError("unsupported scalar field size");
        uint64_t r[4]; detail::read_exact(r_, r, 32);
        if (memcmp(q, detail::FQ_P, 32) || memcmp(r, detail::FR_P, 32)) throw SerializationError("only BN254 zkeys are supported");
        h.n_vars = detail::read_le<uint32_t>(r_); h.n_public = detail::read_le<uint32_t>(r_); h.domain_size = detail::read_le<uint32_t>(r_);
        detail::read_exact(r_, &h.alpha_g1, 64); detail::read_exact(r_, &h.beta_g1, 64);
        detail::read_exact(r_, &h.beta_g2, 128); detail::read_exact(r_, &h.gamma_g2, 128);
        detail::read_exact(r_, &h.delta_g1, 64); detail::read_exact(r_, &h.delta_g2, 128);
        if (h.n_vars < h.n_public + 1) throw SerializationError("bad zkey header");
        return h;
    }

private:
    void seek(uint32_t id) {
        auto it = sections_.find(id);
        if (it == sections_.end()) throw SerializationError("missing zkey section " + std::to_string(id));
        r_.clear(); r_.seekg((std::streamoff)it->second[0].position);
    }
    // points are already Montgomery (zkey.rs:327-332): one bulk read per section instead of per-point byteorder calls.
    // NB the reference checks every point on-curve (G1Affine::new, zkey.rs:347); here b2g_pk_load takes them as given.
    std::vector<G1Affine> g1_section(size_t n, uint32_t id) { seek(id); std::vector<G1Affine> v(n); if (n) detail::read_exact(r_, v.data(), n * 64); return v; }
    std::vector<G2Affine> g2_section(size_t n, uint32_t id) { seek(id); std::vector<G2Affine> v(n); if (n) detail::read_exact(r_, v.data(), n * 128); return v; }

    std::istream& r_;
    std::string ftype_;
    uint32_t version_ = 0;
    std::map<uint32_t, std::vector<detail::Section>> sections_;
};

inline std::pair<ProvingKey, ConstraintMatrices> read_zkey(std::istream& reader) {   // src/zkey.rs:53-60
    BinFile f(reader);
    ProvingKey pk = f.proving_key();
    ConstraintMatrices m = f.matrices();
    return {std::move(pk), std::move(m)};
}

// snarkjs .wtns: "wtns", version, sections {1: n8 u32, prime[n8], nWitness u32; 2: nWitness x n8 canonical LE}
inline std::vector<Fr> read_wtns(std::istream& r) {
    char magic[4]; detail::read_exact(r, magic, 4);
    if (memcmp(magic, "wtns", 4)) throw SerializationError("not a wtns file");
    detail::read_le<uint32_t>(r);
    uint32_t nsec = detail::read_le<uint32_t>(r), nwit = 0;
    std::vector<Fr> out;
    for (uint32_t i = 0; i < nsec; i++) {
        uint32_t id = detail::read_le<uint32_t>(r); uint64_t len = detail::read_le<uint64_t>(r);
        std::streamoff pos = r.tellg();
        if (id == 1) {
            uint32_t n8 = detail::read_le<uint32_t>(r);
            uint64_t prime[4]; if (n8 != 32) throw SerializationError("unsupported field size"); detail::read_exact(r, prime, 32);
            if (memcmp(prime, detail::FR_P, 32)) throw SerializationError("only BN254 witnesses are supported");
            nwit = detail::read_le<uint32_t>(r);
        } else if (id == 2) {
            out.resize(nwit);
            for (uint32_t k = 0; k < nwit; k++) { BigInt256 b; detail::read_exact(r, b.l, 32); out[k] = Fr::from_bigint(b); }
        }
        r.seekg(pos + (std::streamoff)len);
    }
    return out;
}

// ---------------------------------------------------------------------------------------------- device side
// One Gpu = one b2g_ctx (one in-flight proof on one device).  Keys / matrices are uploaded on first use; the handle lives
// in the host object's DeviceSlot (freed with it), never in an address-keyed table.  A Gpu must outlive the proofs issued
// on it, not the keys: b2g_pk_free / b2g_matrices_free need only the device.
class Gpu {
public:
    explicit Gpu(int device = 0) { check(b2g_ctx_create(device, 0, 1, &ctx_)); }
    ~Gpu() { if (ctx_) b2g_ctx_destroy(ctx_); }
    Gpu(const Gpu&) = delete; Gpu& operator=(const Gpu&) = delete;
    static Gpu& instance() { static Gpu g(0); return g; }
    b2g_ctx* ctx() { return ctx_; }

    b2g_pk* pk(const ProvingKey& k) {
        if (void* h = k.device.find(ctx_, 0)) return (b2g_pk*)h;
        b2g_pk_desc d; memset(&d, 0, sizeof d);
        d.n_vars = (uint32_t)k.a_query.size(); d.n_public = (uint32_t)k.vk.gamma_abc_g1.size() - 1; d.domain_size = (uint32_t)k.h_query.size();
        d.alpha_g1 = &k.vk.alpha_g1; d.beta_g1 = &k.beta_g1; d.delta_g1 = &k.delta_g1; d.beta_g2 = &k.vk.beta_g2; d.delta_g2 = &k.vk.delta_g2;
        d.a_query = k.a_query.data(); d.b_g1_query = k.b_g1_query.data(); d.b_g2_query = k.b_g2_query.data();
        d.l_query = k.l_query.data(); d.h_query = k.h_query.data();
        b2g_pk* h = nullptr; check(b2g_pk_load(ctx_, &d, &h));
        k.device.put(ctx_, 0, h, [](void* p) { b2g_pk_free((b2g_pk*)p); });
        return h;
    }

    b2g_mat* mat(const ConstraintMatrices& m, size_t n_vars, uint32_t reduction = B2G_REDUCTION_CIRCOM) {
        const uint32_t tag = reduction | (uint32_t)(n_vars << 1);     // a matrices handle is specific to (reduction, n_vars)
        if (void* h = m.device.find(ctx_, tag)) return (b2g_mat*)h;
        std::vector<uint32_t> rp[3], col[3]; std::vector<Fr> val[3];
        const Matrix* src[3] = {&m.a, &m.b, &m.c};
        const int nmat = reduction == B2G_REDUCTION_LIBSNARK ? 3 : 2;
        for (int k = 0; k < nmat; k++) {
            rp[k].assign(m.num_constraints + 1, 0);
            for (size_t i = 0; i < m.num_constraints; i++) {
                const auto& row = i < src[k]->size() ? (*src[k])[i] : Matrix::value_type();
                for (const auto& e : row) { val[k].push_back(e.first); col[k].push_back((uint32_t)e.second); }
                rp[k][i + 1] = (uint32_t)col[k].size();
            }
        }
        b2g_mat_desc d; memset(&d, 0, sizeof d);
        d.num_constraints = (uint32_t)m.num_constraints; d.num_inputs = (uint32_t)m.num_instance_variables; d.n_vars = (uint32_t)n_vars;
        d.reduction = reduction;
        d.a_rowptr = rp[0].data(); d.a_col = col[0].data(); d.a_val = val[0].data();
        d.b_rowptr = rp[1].data(); d.b_col = col[1].data(); d.b_val = val[1].data();
        if (nmat == 3) { d.c_rowptr = rp[2].data(); d.c_col = col[2].data(); d.c_val = val[2].data(); }
        b2g_mat* h = nullptr; check(b2g_matrices_load(ctx_, &d, &h));
        m.device.put(ctx_, tag, h, [](void* p) { b2g_matrices_free((b2g_mat*)p); });
        return h;
    }

private:
    b2g_ctx* ctx_ = nullptr;
};

template <uint32_t REDUCTION>
struct Reduction {
    static constexpr uint32_t ID = REDUCTION;
    static std::vector<Fr> witness_map_from_matrices(const ConstraintMatrices& matrices, size_t num_inputs, size_t num_constraints,
                                                     const std::vector<Fr>& full_assignment, Gpu& gpu = Gpu::instance()) {
        if (num_inputs != matrices.num_instance_variables || num_constraints != matrices.num_constraints)
            throw SynthesisError("num_inputs / num_constraints disagree with the matrices");
        size_t n = 1; while (n < num_constraints + num_inputs) n <<= 1;
        if (n > (size_t(1) << 27)) throw PolynomialDegreeTooLarge();
        std::vector<Fr> h(n);
        uint32_t dom = 0;
        check(b2g_witness_map(gpu.ctx(), gpu.mat(matrices, full_assignment.size(), REDUCTION), full_assignment.data(), h.data(), &dom));
        return h;
    }
};
typedef Reduction<B2G_REDUCTION_CIRCOM> CircomReduction;       // src/circom/qap.rs:12-14 (snarkjs keys)
typedef Reduction<B2G_REDUCTION_LIBSNARK> LibsnarkReduction;   // ark-groth16's default QAP (tests/groth16.rs:9,25-35; needs matrices.c)

}  // namespace ark_circom
#include "ark_circom_verifier.hpp"
#include "ark_circom_ethereum.hpp"
namespace ark_circom {

template <class QAP = CircomReduction>
struct Groth16T {                                       // Groth16::<Bn254, QAP>
    // verification (host pairing, ark_circom_verifier.hpp): src/zkey.rs:868-870, 914-916; tests/groth16.rs:33-35
    static PreparedVerifyingKey process_vk(const VerifyingKey& vk) { return prepare_verifying_key(vk); }
    static bool verify_with_processed_vk(const PreparedVerifyingKey& pvk, const std::vector<Fr>& public_inputs, const Proof& proof) {
        return ark_circom::verify_with_processed_vk(pvk, public_inputs, proof);
    }
    static bool verify(const VerifyingKey& vk, const std::vector<Fr>& public_inputs, const Proof& proof) {
        return ark_circom::verify_with_processed_vk(prepare_verifying_key(vk), public_inputs, proof);
    }
    static Proof create_proof_with_reduction_and_matrices(const ProvingKey& pk, const Fr& r, const Fr& s, const ConstraintMatrices& matrices,
                                                          size_t num_inputs, size_t num_constraints, const std::vector<Fr>& full_assignment,
                                                          Gpu& gpu = Gpu::instance()) {
        if (num_inputs != matrices.num_instance_variables || num_constraints != matrices.num_constraints)
            throw SynthesisError("num_inputs / num_constraints disagree with the matrices");
        if (full_assignment.size() != pk.a_query.size()) throw SynthesisError("AssignmentMissing: full_assignment length != n_vars");
        BigInt256 rb = r.into_bigint(), sb = s.into_bigint();
        Proof p;
        check(b2g_prove(gpu.ctx(), gpu.pk(pk), gpu.mat(matrices, full_assignment.size(), QAP::ID), rb.l, sb.l, full_assignment.data(), p.bytes));
        return p;
    }

    // Many proofs for one key, `inflight` of them queued on the GPU at any time, driven by THIS thread alone
    // (b2g_prove_submit / b2g_prove_wait on `inflight` contexts): what a rayon pool around the synchronous call does in the
    // reference's world.  rs[i] = (r, s) and assignments[i] = full assignment of proof i; witnesses are page-locked for the
    // duration of the call so that their upload overlaps the previous proofs.  Proofs are returned in order.
    static std::vector<Proof> prove_batch(const ProvingKey& pk, const ConstraintMatrices& matrices, const std::vector<std::pair<Fr, Fr>>& rs,
                                          const std::vector<const std::vector<Fr>*>& assignments, int inflight = 3, int device = 0) {
        if (rs.size() != assignments.size()) throw SynthesisError("prove_batch: one (r, s) per assignment");
        const size_t n = rs.size();
        if (inflight < 1) inflight = 1;
        std::vector<std::unique_ptr<Gpu>> gpus;
        for (int k = 0; k < inflight && (size_t)k < n; k++) gpus.emplace_back(new Gpu(device));
        std::vector<Proof> out(n);
        std::vector<BigInt256> rb(n), sb(n);
        for (size_t i = 0; i < n; i++) {
            if (assignments[i]->size() != pk.a_query.size()) throw SynthesisError("AssignmentMissing: full_assignment length != n_vars");
            rb[i] = rs[i].first.into_bigint(); sb[i] = rs[i].second.into_bigint();
            check(b2g_host_register(assignments[i]->data(), assignments[i]->size() * sizeof(Fr)));
        }
        auto submit = [&](size_t i) {
            Gpu& g = *gpus[i % gpus.size()];
            check(b2g_prove_submit(g.ctx(), g.pk(pk), g.mat(matrices, assignments[i]->size(), QAP::ID), rb[i].l, sb[i].l, assignments[i]->data(), out[i].bytes));
        };
        size_t submitted = 0;
        try {
            for (; submitted < gpus.size(); submitted++) submit(submitted);
            for (size_t done = 0; done < n; done++) {
                check(b2g_prove_wait(gpus[done % gpus.size()]->ctx()));
                if (submitted < n) submit(submitted++);
            }
        } catch (...) {
            for (size_t i = 0; i < n; i++) b2g_host_unregister(assignments[i]->data());
            throw;
        }
        for (size_t i = 0; i < n; i++) b2g_host_unregister(assignments[i]->data());
        return out;
    }

    template <class Rng>
    static Proof prove(const ProvingKey& pk, const ConstraintMatrices& matrices, const std::vector<Fr>& full_assignment, Rng& rng,
                       Gpu& gpu = Gpu::instance()) {
        Fr r = Fr::rand(rng), s = Fr::rand(rng);        // r first, then s (create_random_proof_with_reduction)
        return create_proof_with_reduction_and_matrices(pk, r, s, matrices, matrices.num_instance_variables, matrices.num_constraints,
                                                        full_assignment, gpu);
    }
};

typedef Groth16T<CircomReduction> Groth16;

// ---------------------------------------------------------------------------------------------- R1CS route (host)
// R1CSFile / R1CS: /root/reference/src/circom/r1cs_reader.rs:54-249; to_matrices = what CircomCircuit::generate_constraints
// + ConstraintSystem::to_matrices yield (src/circom/circuit.rs:30-82: column index = wire index).
struct R1CS {
    size_t num_inputs = 0, num_aux = 0, num_variables = 0;
    struct Constraint { std::vector<std::pair<size_t, Fr>> a, b, c; };          // (index, coeff): src/circom/mod.rs:13-14
    std::vector<Constraint> constraints;
    std::vector<uint64_t> wire_mapping;

    static R1CS read(std::istream& r) {
        char magic[4]; detail::read_exact(r, magic, 4);
        if (memcmp(magic, "r1cs", 4)) throw SerializationError("Invalid magic number");
        if (detail::read_le<uint32_t>(r) != 1) throw SerializationError("Unsupported version");
        uint32_t nsec = detail::read_le<uint32_t>(r);
        std::map<uint32_t, std::pair<uint64_t, uint64_t>> sec;
        for (uint32_t i = 0; i < nsec; i++) {
            uint32_t t = detail::read_le<uint32_t>(r); uint64_t sz = detail::read_le<uint64_t>(r);
            sec[t] = {(uint64_t)r.tellg(), sz};
            r.seekg((std::streamoff)sz, std::ios::cur);
        }
        for (uint32_t t : {1u, 2u, 3u}) if (!sec.count(t)) throw SerializationError("missing r1cs section");
        r.clear(); r.seekg((std::streamoff)sec[1].first);
        if (detail::read_le<uint32_t>(r) != 32) throw SerializationError("This parser only supports 32-byte fields");
        uint64_t prime[4]; detail::read_exact(r, prime, 32);
        if (memcmp(prime, detail::FR_P, 32)) throw SerializationError("This parser only supports bn256");
        uint32_t n_wires = detail::read_le<uint32_t>(r), n_pub_out = detail::read_le<uint32_t>(r), n_pub_in = detail::read_le<uint32_t>(r);
        detail::read_le<uint32_t>(r); detail::read_le<uint64_t>(r);
        uint32_t n_cons = detail::read_le<uint32_t>(r);
        R1CS out;
        out.num_inputs = 1 + n_pub_in + n_pub_out; out.num_variables = n_wires; out.num_aux = n_wires - out.num_inputs;
        r.seekg((std::streamoff)sec[2].first);
        auto lc = [&](std::vector<std::pair<size_t, Fr>>& v) {
            uint32_t n = detail::read_le<uint32_t>(r);
            for (uint32_t k = 0; k < n; k++) { uint32_t w = detail::read_le<uint32_t>(r); BigInt256 b; detail::read_exact(r, b.l, 32); v.push_back({w, Fr::from_bigint(b)}); }
        };
        out.constraints.resize(n_cons);
        for (auto& c : out.constraints) { lc(c.a); lc(c.b); lc(c.c); }
        if (sec[3].second != (uint64_t)n_wires * 8) throw SerializationError("Invalid map section size");
        r.seekg((std::streamoff)sec[3].first);
        out.wire_mapping.resize(n_wires);
        if (n_wires) detail::read_exact(r, out.wire_mapping.data(), (size_t)n_wires * 8);
        if (n_wires && out.wire_mapping[0] != 0) throw SerializationError("Wire 0 should always be mapped to 0");
        return out;
    }

    ConstraintMatrices to_matrices() const {
        ConstraintMatrices m;
        m.num_instance_variables = num_inputs; m.num_witness_variables = num_aux; m.num_constraints = constraints.size();
        m.a.resize(constraints.size()); m.b.resize(constraints.size()); m.c.resize(constraints.size());
        for (size_t i = 0; i < constraints.size(); i++) {
            for (auto& e : constraints[i].a) m.a[i].push_back({e.second, e.first});
            for (auto& e : constraints[i].b) m.b[i].push_back({e.second, e.first});
            for (auto& e : constraints[i].c) m.c[i].push_back({e.second, e.first});
            m.a_num_non_zero += m.a[i].size(); m.b_num_non_zero += m.b[i].size(); m.c_num_non_zero += m.c[i].size();
        }
        return m;
    }
};

}  // namespace ark_circom
