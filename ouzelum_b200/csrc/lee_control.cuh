// Lee geometric controllers, one env per thread, all in registers (kernel K4).  Every TU is compiled with -fmad=false and the
// multiply-adds below are explicit (fmaf), so every kernel that uses these functions produces identical bits.
// CPU twin: oracle/lee_control.py (same component formulas).  Reference:
//   isaacgymenvs/controllers/position_control.py:19-109, velocity_control.py:17-112, attitude_control.py:17-78,
//   rotation_conversions.py:36-64,149-171,216-255, math_control.py:10-16, controller.py:45-48.
// The reference issues ~60 batched torch launches (6 bmm of 3x3) per call; this is one launch, 68 B in + 16 B out per env.
#pragma once
#include <math.h>

namespace ozl {

struct LeeGains {
    float kP[3], kV[3], kR[3], kO[3], scale[4];
};
enum { LEE_POSITION = 0, LEE_VELOCITY = 1, LEE_ATTITUDE = 2 };

struct M3 { float m[3][3]; };

__device__ __forceinline__ M3 mm3(const M3& A, const M3& B) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C.m[i][j] = fmaf(A.m[i][2], B.m[2][j], fmaf(A.m[i][1], B.m[1][j], A.m[i][0] * B.m[0][j]));
    return C;
}
__device__ __forceinline__ M3 tr3(const M3& A) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[j][i];
    return C;
}
__device__ __forceinline__ void mv3(const M3& A, const float v[3], float o[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = fmaf(A.m[i][2], v[2], fmaf(A.m[i][1], v[1], A.m[i][0] * v[0]));
}
__device__ __forceinline__ void cr3(const float a[3], const float b[3], float o[3]) {
    o[0] = fmaf(a[1], b[2], -(a[2] * b[1]));
    o[1] = fmaf(a[2], b[0], -(a[0] * b[2]));
    o[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}
// euler_angles_to_matrix((yaw, pitch, roll), "ZYX") = Rz Ry Rx   (rotation_conversions.py:149-171)
__device__ __forceinline__ M3 R_from_zyx(float yaw, float pitch, float roll) {
    float sz, cz, sy, cy, sx, cx;
    sincosf(yaw, &sz, &cz); sincosf(pitch, &sy, &cy); sincosf(roll, &sx, &cx);
    M3 Rz = {{{cz, -sz, 0.f}, {sz, cz, 0.f}, {0.f, 0.f, 1.f}}};
    M3 Ry = {{{cy, 0.f, sy}, {0.f, 1.f, 0.f}, {-sy, 0.f, cy}}};
    M3 Rx = {{{1.f, 0.f, 0.f}, {0.f, cx, -sx}, {0.f, sx, cx}}};
    return mm3(mm3(Rz, Ry), Rx);
}

// state: p[3], q xyzw[4], v[3], w[3] (world).  cmd: 4 floats (already multiplied by scale_input).
__device__ __forceinline__ void lee_control(int mode, const float p[3], const float q[4], const float v[3], const float w[3],
                                            const float cmd[4], const LeeGains& g, float& thrust, float torque[3]) {
    const float r = q[3], i = q[0], j = q[1], k = q[2];                       // state[:, [6,3,4,5]] -> wxyz
    const float two_s = 2.0f / fmaf(k, k, fmaf(j, j, fmaf(i, i, r * r)));           // rotation_conversions.py:50
    M3 R = {{{fmaf(-two_s, fmaf(k, k, j * j), 1.f), two_s * fmaf(i, j, -(k * r)), two_s * fmaf(i, k, j * r)},
             {two_s * fmaf(i, j, k * r), fmaf(-two_s, fmaf(k, k, i * i), 1.f), two_s * fmaf(j, k, -(i * r))},
             {two_s * fmaf(i, k, -(j * r)), two_s * fmaf(j, k, i * r), fmaf(-two_s, fmaf(j, j, i * i), 1.f)}}};
    // matrix_to_euler_angles(R, "ZYX")[:, [2,1,0]] -> roll, pitch, yaw
    const float yaw = atan2f(R.m[1][0], R.m[0][0]);
    const float pitch = asinf(-R.m[2][0]);
    const float roll = atan2f(R.m[2][1], R.m[2][2]);
    M3 Rd;
    float yaw_rate;
    if (mode == LEE_POSITION) {
        float acc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) acc[a] = fmaf(g.kP[a], cmd[a] - p[a], -(g.kV[a] * v[a]));     // position_control.py:39-40
        acc[2] += 1.0f;
        thrust = fmaf(acc[2], R.m[2][2], fmaf(acc[1], R.m[1][2], acc[0] * R.m[0][2]));              // :44
        const float n = sqrtf(fmaf(acc[2], acc[2], fmaf(acc[1], acc[1], acc[0] * acc[0])));
        const float b3[3] = {acc[0] / n, acc[1] / n, acc[2] / n};
        float sy_, cy_;
        sincosf(yaw, &sy_, &cy_);
        const float tmp[3] = {cy_, sy_, 0.0f};
        float b2[3], b1[3];
        cr3(b3, tmp, b2);
        const float n2 = sqrtf(fmaf(b2[2], b2[2], fmaf(b2[1], b2[1], b2[0] * b2[0])));
        b2[0] /= n2; b2[1] /= n2; b2[2] /= n2;
        cr3(b2, b3, b1);
#pragma unroll
        for (int a = 0; a < 3; ++a) { Rd.m[a][0] = b1[a]; Rd.m[a][1] = b2[a]; Rd.m[a][2] = b3[a]; }
        const float two_pi = 3.14159265358979323846f * 2.0f;
        float yr = fmodf(cmd[3] - yaw, two_pi);                                 // torch.remainder: sign of the divisor
        if (yr < 0.0f) yr += two_pi;
        yaw_rate = (yr > 3.14159265358979323846f) ? (yr - two_pi) : yr;          // :90-92
    } else if (mode == LEE_VELOCITY) {
        const M3 Rv = R_from_zyx(yaw, 0.0f, 0.0f);                               // velocity_control.py:32-38
        float vv[3];
        mv3(tr3(Rv), v, vv);
        float acc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) acc[a] = g.kV[a] * (cmd[a] - vv[a]);         // :46-47
        acc[2] += 1.0f;
        thrust = fmaf(acc[2], R.m[2][2], fmaf(acc[1], R.m[1][2], acc[0] * R.m[0][2]));
        const float pitch_sp = atan2f(acc[0], acc[2]);                           // :58
        const float roll_sp = atan2f(-acc[1], sqrtf(acc[2] * acc[2] + acc[0] * acc[0]));   // :59-60
        Rd = R_from_zyx(yaw, pitch_sp, roll_sp);
        yaw_rate = cmd[3];
    } else {
        Rd = R_from_zyx(yaw, cmd[2], cmd[1]);                                    // attitude_control.py:33-35,58-59
        thrust = cmd[0] + 1.0f;                                                  // :78
        yaw_rate = cmd[3];
    }
    // rotation error, body-rate setpoint, torque (position_control.py:66-109)
    const M3 Rt = tr3(R);
    const M3 A = mm3(tr3(Rd), R), B = mm3(Rt, Rd);
    const float e_R[3] = {0.5f * -(A.m[1][2] - B.m[1][2]), 0.5f * (A.m[0][2] - B.m[0][2]), 0.5f * -(A.m[0][1] - B.m[0][1])};
    float sp, cp, sr, cr;
    sincosf(pitch, &sp, &cp);
    sincosf(roll, &sr, &cr);
    const float wdb[3] = {-sp * yaw_rate, (sr * cp) * yaw_rate, (cr * cp) * yaw_rate};
    float t1[3], desired[3], actual[3];
    mv3(Rd, wdb, t1);
    mv3(Rt, t1, desired);
    mv3(Rt, w, actual);
#pragma unroll
    for (int a = 0; a < 3; ++a) torque[a] = fmaf(-g.kR[a], e_R[a], -(g.kO[a] * (actual[a] - desired[a])));
}

}  // namespace ozl
