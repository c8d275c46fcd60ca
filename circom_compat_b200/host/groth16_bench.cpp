// groth16_bench.cpp - C++ counterpart of /root/reference/benches/groth16.rs:13-85: read a zkey, obtain the full assignment,
// draw / take (r, s), call Groth16::create_proof_with_reduction_and_matrices repeatedly and report the time per proof.
//
//   groth16_bench --parse-only <circuit.zkey> [--dump-key]         host-only: print what read_zkey produced (no GPU)
//   groth16_bench --verify <circuit.zkey> <proof_hex> [inputs...]  host-only: process_vk + verify_with_processed_vk
//   groth16_bench --ethereum <circuit.zkey> <proof_hex> [inputs...] host-only: src/ethereum.rs views of vk / proof / inputs, both directions
//   groth16_bench <circuit.zkey> chain:<a>|<witness.wtns> [iters] [r_hex s_hex]
//       chain:<a> = the witness of the reference's squaring-chain bench family for input a
//       (test-vectors/complex-circuit/input.json has a = 3), computed on the host instead of by WASM.
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "ark_circom_b200.hpp"

using namespace ark_circom;

static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ULL) {
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; }
    return h;
}

static BigInt256 parse_hex(const std::string& s) {
    BigInt256 b = {{0, 0, 0, 0}};
    std::string t = s.rfind("0x", 0) == 0 ? s.substr(2) : s;
    if (t.size() > 64) throw std::invalid_argument("scalar too long");
    for (char c : t) {
        int v = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1;
        if (v < 0) throw std::invalid_argument("bad hex digit");
        for (int i = 3; i > 0; i--) b.l[i] = (b.l[i] << 4) | (b.l[i - 1] >> 60);
        b.l[0] = (b.l[0] << 4) | (uint64_t)v;
    }
    return b;
}

// witness of the squaring chain [1, c, a, a^2, a^4, ...] (App. B.4 of SURVEY.md), Montgomery form
static std::vector<Fr> chain_witness(size_t n_vars, uint64_t a) {
    std::vector<Fr> w(n_vars);
    w[0] = Fr::from_u64(1);
    if (n_vars > 2) w[2] = Fr::from_u64(a);
    for (size_t k = 3; k < n_vars; k++) detail::fr_mont_mul(w[k].l, w[k - 1].l, w[k - 1].l);
    if (n_vars > 2) detail::fr_mont_mul(w[1].l, w[n_vars - 1].l, w[n_vars - 1].l);
    return w;
}

int main(int argc, char** argv) {
    try {
        if (argc >= 3 && std::string(argv[1]) == "--parse-only") {
            std::ifstream f(argv[2], std::ios::binary);
            if (!f) throw SerializationError("cannot open zkey");
            auto kv = read_zkey(f);
            const ProvingKey& pk = kv.first; const ConstraintMatrices& m = kv.second;
            std::printf("n_vars=%zu n_public=%zu domain=%zu num_constraints=%zu num_instance=%zu num_witness=%zu a_nnz=%zu b_nnz=%zu\n",
                        pk.a_query.size(), pk.vk.gamma_abc_g1.size() - 1, pk.h_query.size(), m.num_constraints, m.num_instance_variables,
                        m.num_witness_variables, m.a_num_non_zero, m.b_num_non_zero);
            std::printf("fnv a=%016llx b1=%016llx b2=%016llx l=%016llx h=%016llx alpha=%016llx\n",
                        (unsigned long long)fnv(pk.a_query.data(), pk.a_query.size() * 64), (unsigned long long)fnv(pk.b_g1_query.data(), pk.b_g1_query.size() * 64),
                        (unsigned long long)fnv(pk.b_g2_query.data(), pk.b_g2_query.size() * 128), (unsigned long long)fnv(pk.l_query.data(), pk.l_query.size() * 64),
                        (unsigned long long)fnv(pk.h_query.data(), pk.h_query.size() * 64), (unsigned long long)fnv(&pk.vk.alpha_g1, 64));
            uint64_t hc = 1469598103934665603ULL;
            for (const Matrix* mm : {&m.a, &m.b})
                for (const auto& row : *mm) for (const auto& e : row) { hc = fnv(e.first.l, 32, hc); uint32_t c = (uint32_t)e.second; hc = fnv(&c, 4, hc); }
            std::printf("fnv coefs=%016llx\n", (unsigned long long)hc);
            if (argc >= 4 && std::string(argv[3]) == "--dump-key") {       // every query point as the zkey's own bytes (small keys)
                auto dump = [](const char* name, const void* p, size_t count, size_t stride) {
                    const uint8_t* b = (const uint8_t*)p;
                    for (size_t i = 0; i < count; i++) {
                        std::printf("%s[%zu]=", name, i);
                        for (size_t k = 0; k < stride; k++) std::printf("%02x", b[i * stride + k]);
                        std::printf("\n");
                    }
                };
                dump("gamma_abc_g1", pk.vk.gamma_abc_g1.data(), pk.vk.gamma_abc_g1.size(), 64);
                dump("a_query", pk.a_query.data(), pk.a_query.size(), 64);
                dump("b_g1_query", pk.b_g1_query.data(), pk.b_g1_query.size(), 64);
                dump("b_g2_query", pk.b_g2_query.data(), pk.b_g2_query.size(), 128);
                dump("l_query", pk.l_query.data(), pk.l_query.size(), 64);
                dump("h_query", pk.h_query.data(), pk.h_query.size(), 64);
            }
            return 0;
        }
        if (argc >= 4 && std::string(argv[1]) == "--verify") {              // host-only: --verify <zkey> <proof_hex> [public inputs, decimal u64 or 0x hex]
            std::ifstream f(argv[2], std::ios::binary);
            if (!f) throw SerializationError("cannot open zkey");
            auto kv = read_zkey(f);
            std::string hx = argv[3];
            if (hx.size() != 512) throw std::invalid_argument("proof must be 256 bytes of hex");
            Proof proof;
            for (int i = 0; i < 256; i++) proof.bytes[i] = (uint8_t)std::stoul(hx.substr(2 * i, 2), nullptr, 16);
            std::vector<Fr> inputs;
            for (int i = 4; i < argc; i++) { std::string a = argv[i]; inputs.push_back(a.rfind("0x", 0) == 0 ? Fr::from_bigint(parse_hex(a)) : Fr::from_u64(std::stoull(a))); }
            auto pvk = Groth16::process_vk(kv.first.vk);                    // src/zkey.rs:868
            std::printf("verified=%d\n", Groth16::verify_with_processed_vk(pvk, inputs, proof) ? 1 : 0);
            return 0;
        }
        if (argc >= 4 && std::string(argv[1]) == "--ethereum") {            // host-only: the Ethereum views (src/ethereum.rs) and their way back
            namespace eth = ark_circom::ethereum;
            std::ifstream f(argv[2], std::ios::binary);
            if (!f) throw SerializationError("cannot open zkey");
            auto kv = read_zkey(f);
            std::string hx = argv[3];
            if (hx.size() != 512) throw std::invalid_argument("proof must be 256 bytes of hex");
            Proof proof;
            for (int i = 0; i < 256; i++) proof.bytes[i] = (uint8_t)std::stoul(hx.substr(2 * i, 2), nullptr, 16);
            std::vector<Fr> inputs;
            for (int i = 4; i < argc; i++) { std::string a = argv[i]; inputs.push_back(a.rfind("0x", 0) == 0 ? Fr::from_bigint(parse_hex(a)) : Fr::from_u64(std::stoull(a))); }
            const eth::VerifyingKey evk = eth::VerifyingKey::from(kv.first.vk);
            const eth::Proof ep = eth::Proof::from(proof);
            auto g1s = [](const eth::G1& g) { auto t = g.as_tuple(); return t[0].hex() + "," + t[1].hex(); };
            auto g2s = [](const eth::G2& g) { auto t = g.as_tuple(); return t[0][0].hex() + "," + t[0][1].hex() + "," + t[1][0].hex() + "," + t[1][1].hex(); };
            std::printf("vk.alpha1=%s\nvk.beta2=%s\nvk.gamma2=%s\nvk.delta2=%s\n", g1s(evk.alpha1).c_str(), g2s(evk.beta2).c_str(), g2s(evk.gamma2).c_str(), g2s(evk.delta2).c_str());
            for (size_t i = 0; i < evk.ic.size(); i++) std::printf("vk.ic[%zu]=%s\n", i, g1s(evk.ic[i]).c_str());
            std::printf("proof.a=%s\nproof.b=%s\nproof.c=%s\ncalldata=", g1s(ep.a).c_str(), g2s(ep.b).c_str(), g1s(ep.c).c_str());
            for (const eth::U256& w : ep.calldata_words()) std::printf("%s", w.hex().c_str());
            std::printf("\n");
            const std::vector<eth::U256> ein = eth::inputs(inputs);
            for (size_t i = 0; i < ein.size(); i++) std::printf("inputs[%zu]=%s\n", i, ein[i].hex().c_str());
            // the reference's convert_vk / convert_proof / convert_fr tests (src/ethereum.rs:195-279), then check_proof with the host verifier
            const VerifyingKey vk2 = evk.into();
            bool rt = !memcmp(&vk2.alpha_g1, &kv.first.vk.alpha_g1, 64) && !memcmp(&vk2.beta_g2, &kv.first.vk.beta_g2, 128) &&
                      !memcmp(&vk2.gamma_g2, &kv.first.vk.gamma_g2, 128) && !memcmp(&vk2.delta_g2, &kv.first.vk.delta_g2, 128) &&
                      vk2.gamma_abc_g1.size() == kv.first.vk.gamma_abc_g1.size();
            for (size_t i = 0; rt && i < vk2.gamma_abc_g1.size(); i++) rt = !memcmp(&vk2.gamma_abc_g1[i], &kv.first.vk.gamma_abc_g1[i], 64);
            const Proof p2 = ep.into();
            rt = rt && !memcmp(p2.bytes, proof.bytes, 256) && eth::Proof::from(p2) == ep;
            std::vector<Fr> in2;
            for (const eth::U256& w : ein) in2.push_back(eth::u256_to_fr(w));
            for (size_t i = 0; i < inputs.size(); i++) rt = rt && in2[i] == inputs[i];
            std::printf("roundtrip=%d\n", rt ? 1 : 0);
            std::printf("verified=%d\n", Groth16::verify_with_processed_vk(Groth16::process_vk(vk2), in2, p2) ? 1 : 0);
            return 0;
        }
        if (argc < 3) { std::fprintf(stderr, "usage: %s [--parse-only] <zkey> chain:<a>|<wtns> [iters] [r_hex s_hex]\n", argv[0]); return 2; }
        std::ifstream f(argv[1], std::ios::binary);
        if (!f) throw SerializationError("cannot open zkey");
        auto kv = read_zkey(f);                                     // benches/groth16.rs:20-23
        const ProvingKey& params = kv.first; const ConstraintMatrices& matrices = kv.second;
        const size_t num_inputs = matrices.num_instance_variables, num_constraints = matrices.num_constraints;
        std::string wsrc = argv[2];
        std::vector<Fr> full_assignment;
        if (wsrc.rfind("chain:", 0) == 0) full_assignment = chain_witness(params.a_query.size(), std::stoull(wsrc.substr(6)));
        else { std::ifstream wf(wsrc, std::ios::binary); if (!wf) throw SerializationError("cannot open wtns"); full_assignment = read_wtns(wf); }
        int iters = argc > 3 ? std::atoi(argv[3]) : 10;
        Fr r, s;
        if (argc > 5) { r = Fr::from_bigint(parse_hex(argv[4])); s = Fr::from_bigint(parse_hex(argv[5])); }
        else { std::mt19937_64 rng(0xB200); r = Fr::rand(rng); s = Fr::rand(rng); }      // benches/groth16.rs:45-50
        Proof proof = Groth16::create_proof_with_reduction_and_matrices(params, r, s, matrices, num_inputs, num_constraints, full_assignment);
        std::printf("proof=%s\n", proof.hex().c_str());
        {   // src/zkey.rs:868-872: process_vk, public inputs = w[1..num_inputs] (circuit.rs:18-26), verify_with_processed_vk
            auto pvk = Groth16::process_vk(params.vk);
            std::vector<Fr> inputs(full_assignment.begin() + 1, full_assignment.begin() + num_inputs);
            std::printf("verified=%d\n", Groth16::verify_with_processed_vk(pvk, inputs, proof) ? 1 : 0);
        }
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < iters; i++)                             // benches/groth16.rs:69-84
            proof = Groth16::create_proof_with_reduction_and_matrices(params, r, s, matrices, num_inputs, num_constraints, full_assignment);
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (iters > 0 ? iters : 1);
        std::printf("groth proof %zu constraints: %.3f ms/proof over %d iterations\n", num_constraints, ms, iters);
        if (const char* inf = std::getenv("B2G_INFLIGHT")) {              // pipelined: one host thread, several proofs queued on the GPU
            const int k = std::atoi(inf), total = iters > 0 ? 3 * iters : 3;
            std::vector<std::pair<Fr, Fr>> rs((size_t)total, {r, s});
            std::vector<const std::vector<Fr>*> ws((size_t)total, &full_assignment);
            Groth16::prove_batch(params, matrices, {rs[0]}, {ws[0]}, 1);                        // warm-up: loads the key on a fresh context
            auto t1 = std::chrono::steady_clock::now();
            std::vector<Proof> proofs = Groth16::prove_batch(params, matrices, rs, ws, k);
            double pms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count() / total;
            bool same = true;
            for (const Proof& q : proofs) same = same && !memcmp(q.bytes, proof.bytes, 256);
            std::printf("pipelined (%d in flight, one host thread, includes key load on %d contexts): %.3f ms/proof over %d proofs, identical=%d\n", k, k, pms, total, same ? 1 : 0);
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
