// Host-side weight preparation for the ConvolutionalNet kernels: BatchNorm folding and
// re-layout of a reference-format state_dict (reference deep_mcts/convolutionalnet.py:22-170;
// key names in SURVEY.md 8 a-N) into the layouts the kernels read.
#pragma once
#include <cuda_bf16.h>

#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "games.cuh"

constexpr float kBnEps = 1e-5f;  // nn.BatchNorm2d default (convolutionalnet.py:82)

// Geometry of the padded flat row space used by the bf16 implicit-GEMM trunk: board b, cell
// (y, x) lives in row  halo + b*P + y*Wp + x  with Wp = g+1 and P = (g+1)^2; every other row
// is a zero pad, so a 3x3 tap (dy, dx) is a plain row shift of dy*Wp + dx.
struct RowGeom {
    int g, Wp, P, halo;
    __host__ __device__ static RowGeom make(int g) {
        RowGeom r;
        r.g = g;
        r.Wp = g + 1;
        r.P = (g + 1) * (g + 1);
        r.halo = g + 2;
        return r;
    }
};

struct HostWeights {
    int C, R, H, g, A;
    // folded fp32 parameters
    std::vector<float> conv1_w;               // [tap 9][cin 3][C]
    std::vector<float> conv1_b;               // [C]
    std::vector<std::vector<float>> trunk_w;  // 2R+1 layers (last = policy head conv): [tap][cin][cout]
    std::vector<std::vector<float>> trunk_b;  // [C]
    std::vector<float> vconv_w;               // [C]
    float vconv_b;
    std::vector<float> vfc1_w, vfc1_b, vfc2_w;  // [H][g*g], [H], [H]
    float vfc2_b;
    std::vector<float> pfc_w;  // [K = g*g*C (k = pos*C + c)][NA2]  (NA2 = A rounded up to even)
    std::vector<float> pfc_b;  // [A]
    int NA2;
};

inline const std::vector<float>* find_tensor(const std::map<std::string, std::vector<float>>& t,
                                             const std::string& key, size_t numel, std::string& err) {
    auto it = t.find(key);
    if (it == t.end()) {
        err = "missing state_dict tensor: " + key;
        return nullptr;
    }
    if (it->second.size() != numel) {
        err = "state_dict tensor " + key + " has " + std::to_string(it->second.size()) + " elements, expected " +
              std::to_string(numel);
        return nullptr;
    }
    return &it->second;
}

// conv [cout][cin][k][k] + BN(cout) -> w'[tap][cin][cout], b'[cout]
inline bool fold_conv_bn(const std::map<std::string, std::vector<float>>& t, const std::string& conv,
                         const std::string& bn, int cout, int cin, int k, std::vector<float>& w,
                         std::vector<float>& b, std::string& err) {
    auto W = find_tensor(t, conv + ".weight", (size_t)cout * cin * k * k, err);
    auto B = W ? find_tensor(t, conv + ".bias", cout, err) : nullptr;
    auto G = B ? find_tensor(t, bn + ".weight", cout, err) : nullptr;
    auto Be = G ? find_tensor(t, bn + ".bias", cout, err) : nullptr;
    auto Mu = Be ? find_tensor(t, bn + ".running_mean", cout, err) : nullptr;
    auto Var = Mu ? find_tensor(t, bn + ".running_var", cout, err) : nullptr;
    if (!Var) return false;
    w.assign((size_t)k * k * cin * cout, 0.f);
    b.assign(cout, 0.f);
    for (int co = 0; co < cout; ++co) {
        float scale = (*G)[co] / std::sqrt((*Var)[co] + kBnEps);
        b[co] = ((*B)[co] - (*Mu)[co]) * scale + (*Be)[co];
        for (int ci = 0; ci < cin; ++ci)
            for (int tap = 0; tap < k * k; ++tap)
                w[((size_t)tap * cin + ci) * cout + co] = (*W)[((size_t)co * cin + ci) * k * k + tap] * scale;
    }
    return true;
}

inline bool build_host_weights(const std::map<std::string, std::vector<float>>& t, int C, int R, int H, int g, int A,
                               HostWeights& hw, std::string& err) {
    hw.C = C; hw.R = R; hw.H = H; hw.g = g; hw.A = A;
    if (!fold_conv_bn(t, "conv1.conv1", "conv1.bn1", C, 3, 3, hw.conv1_w, hw.conv1_b, err)) return false;
    hw.trunk_w.assign(2 * R + 1, {});
    hw.trunk_b.assign(2 * R + 1, {});
    for (int i = 0; i < R; ++i) {
        std::string p = "residual_blocks." + std::to_string(i) + ".";
        if (!fold_conv_bn(t, p + "conv1", p + "bn1", C, C, 3, hw.trunk_w[2 * i], hw.trunk_b[2 * i], err)) return false;
        if (!fold_conv_bn(t, p + "conv2", p + "bn2", C, C, 3, hw.trunk_w[2 * i + 1], hw.trunk_b[2 * i + 1], err))
            return false;
    }
    if (!fold_conv_bn(t, "policy_head.conv1.conv1", "policy_head.conv1.bn1", C, C, 3, hw.trunk_w[2 * R],
                      hw.trunk_b[2 * R], err))
        return false;
    std::vector<float> vb;
    if (!fold_conv_bn(t, "value_head.conv1.conv1", "value_head.conv1.bn1", 1, C, 1, hw.vconv_w, vb, err)) return false;
    hw.vconv_b = vb[0];
    int P = g * g;
    auto W1 = find_tensor(t, "value_head.fc1.weight", (size_t)H * P, err);
    auto B1 = W1 ? find_tensor(t, "value_head.fc1.bias", H, err) : nullptr;
    auto W2 = B1 ? find_tensor(t, "value_head.fc2.weight", H, err) : nullptr;
    auto B2 = W2 ? find_tensor(t, "value_head.fc2.bias", 1, err) : nullptr;
    auto WP = B2 ? find_tensor(t, "policy_head.fc1.weight", (size_t)A * C * P, err) : nullptr;
    auto BP = WP ? find_tensor(t, "policy_head.fc1.bias", A, err) : nullptr;
    if (!BP) return false;
    hw.vfc1_w = *W1; hw.vfc1_b = *B1; hw.vfc2_w = *W2; hw.vfc2_b = (*B2)[0];
    hw.pfc_b = *BP;
    hw.NA2 = (A + 1) & ~1;
    hw.pfc_w.assign((size_t)P * C * hw.NA2, 0.f);
    // reference flatten order is (c, y, x): W[a][c*P + pos]  ->  W''[pos*C + c][a]
    for (int a = 0; a < A; ++a)
        for (int c = 0; c < C; ++c)
            for (int pos = 0; pos < P; ++pos)
                hw.pfc_w[((size_t)pos * C + c) * hw.NA2 + a] = (*WP)[(size_t)a * C * P + (size_t)c * P + pos];
    return true;
}

// bf16 image of one trunk layer exactly as it sits in shared memory for tcgen05.mma:
// per tap a K-major, no-swizzle B operand  [kc = cin/8][n = cout][8 cin]  (32 KB at C = 128).
inline void pack_trunk_bf16(const std::vector<float>& w, int C, std::vector<__nv_bfloat16>& out) {
    out.resize((size_t)9 * C * C);
    for (int tap = 0; tap < 9; ++tap)
        for (int ci = 0; ci < C; ++ci)
            for (int co = 0; co < C; ++co)
                out[(((size_t)tap * (C / 8) + ci / 8) * C + co) * 8 + (ci % 8)] =
                    __float2bfloat16(w[((size_t)tap * C + ci) * C + co]);
}
