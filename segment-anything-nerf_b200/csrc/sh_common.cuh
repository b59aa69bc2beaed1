// Real spherical harmonics by recurrence (shared by the SH encoder kernel and the fused view head).
// See encoders_misc.cu for the derivation and the reference lines it replaces (shencoder.cu:27-355).
#pragma once

#include "common.cuh"

namespace sanerf {

// N_l^m = sqrt((2l+1)/(4 pi) * (l-m)!/(l+m)!) * (m ? sqrt(2)*(-1)^m : 1), l < 8, m <= l
// (values printed to 17 digits by the formula above; tests/test_oracle_sh.py re-derives them)
static __device__ __constant__ const float kShNorm[8][8] = {
    {0.28209479177387814f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f},
    {0.48860251190291992f, -0.48860251190291998f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f},
    {0.63078313050504009f, -0.36418281019735976f, 0.18209140509867988f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f},
    {0.7463526651802308f, -0.3046971996429772f, 0.096353714754685155f, -0.039336239328442907f, 0.0f, 0.0f, 0.0f, 0.0f},
    {0.84628437532163447f, -0.26761861742291571f, 0.063078313050504001f, -0.016858388283618388f, 0.0059603403376112026f, 0.0f, 0.0f, 0.0f},
    {0.9356025796273888f, -0.24157154730437169f, 0.045652731285460234f, -0.0093188247511476283f, 0.0021964680580751762f, -0.00069458418713245519f, 0.0f, 0.0f},
    {1.0171072362820548f, -0.22195099524523101f, 0.03509353369580661f, -0.0058489222826344353f, 0.0010678622237644956f, -0.00022766899107568562f, 6.5722376641838803e-05f, 0.0f},
    {1.0925484305920792f, -0.20647224590289676f, 0.028097313806030647f, -0.0039735602250741348f, 0.00059903674311141165f, -9.9839457185235285e-05f, 1.9580128477462541e-05f, -5.233009453691466e-06f}};

template <int DEG, bool JAC>
__device__ __forceinline__ void sh_eval(float x, float y, float z, float* __restrict__ out,
                                        float* __restrict__ jac /* [3][DEG*DEG] */) {
    constexpr int C2 = DEG * DEG;
    // (x + i y)^m
    float A[DEG], Bm[DEG];
    A[0] = 1.0f; Bm[0] = 0.0f;
#pragma unroll
    for (int m = 1; m < DEG; ++m) {
        A[m] = x * A[m - 1] - y * Bm[m - 1];
        Bm[m] = x * Bm[m - 1] + y * A[m - 1];
    }
    // Q[l][m] = d^m/dz^m P_l(z); one extra column so that Q[l][l+1] = 0 is addressable
    float Q[DEG][DEG + 1];
#pragma unroll
    for (int l = 0; l < DEG; ++l)
#pragma unroll
        for (int m = 0; m <= DEG; ++m) Q[l][m] = 0.0f;
    float dfact = 1.0f;  // (2m-1)!!
#pragma unroll
    for (int m = 0; m < DEG; ++m) {
        if (m > 0) dfact *= (float)(2 * m - 1);
        Q[m][m] = dfact;
        if (m + 1 < DEG) Q[m + 1][m] = (float)(2 * m + 1) * z * dfact;
#pragma unroll
        for (int l = m + 2; l < DEG; ++l)
            Q[l][m] = ((float)(2 * l - 1) * z * Q[l - 1][m] - (float)(l + m - 1) * Q[l - 2][m]) *
                      (1.0f / (float)(l - m));
    }
#pragma unroll
    for (int l = 0; l < DEG; ++l) {
        const int centre = l * l + l;
        // m = 0
        out[centre] = kShNorm[l][0] * Q[l][0];
        if (JAC) {
            jac[0 * C2 + centre] = 0.0f;
            jac[1 * C2 + centre] = 0.0f;
            jac[2 * C2 + centre] = kShNorm[l][0] * Q[l][1];
        }
#pragma unroll
        for (int m = 1; m <= l; ++m) {
            const float nq = kShNorm[l][m] * Q[l][m];
            out[centre + m] = nq * A[m];
            out[centre - m] = nq * Bm[m];
            if (JAC) {
                const float nq1 = kShNorm[l][m] * Q[l][m + 1];
                const float fm = (float)m;
                jac[0 * C2 + centre + m] = nq * fm * A[m - 1];
                jac[1 * C2 + centre + m] = -nq * fm * Bm[m - 1];
                jac[2 * C2 + centre + m] = nq1 * A[m];
                jac[0 * C2 + centre - m] = nq * fm * Bm[m - 1];
                jac[1 * C2 + centre - m] = nq * fm * A[m - 1];
                jac[2 * C2 + centre - m] = nq1 * Bm[m];
            }
        }
    }
}

}  // namespace sanerf
