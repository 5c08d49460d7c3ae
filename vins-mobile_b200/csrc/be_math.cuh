// be_math.cuh -- f64 device helpers for the back end: 3-vectors, 3x3 row-major matrices, quaternions stored (x,y,z,w)
// like Eigen::Quaterniond::coeffs() / para_Pose[3..6] (VINS.cpp:97-101), and the Utility:: helpers the factors use
// (/root/reference/VINS_ios/utility.hpp:21-121).
#pragma once
#include "common.cuh"

namespace be {

struct V3 { double x, y, z; };
struct Q4 { double x, y, z, w; };
struct M3 { double m[9]; };      // row-major

__device__ __forceinline__ V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 ld3(const double *p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(double *p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double norm(V3 a) { return sqrt(dot(a, a)); }

__device__ __forceinline__ M3 ldm(const double *p) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = p[i]; return r; }
__device__ __forceinline__ void stm(double *p, const M3 &a) { for (int i = 0; i < 9; i++) p[i] = a.m[i]; }
__device__ __forceinline__ M3 eye3() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = (i % 4 == 0) ? 1.0 : 0.0; return r; }
__device__ __forceinline__ V3 operator*(const M3 &a, V3 v) {
    return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
__device__ __forceinline__ M3 operator*(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
    return r;
}
__device__ __forceinline__ M3 operator*(double s, const M3 &a) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = s * a.m[i]; return r; }
__device__ __forceinline__ M3 operator+(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i]; return r; }
__device__ __forceinline__ M3 operator-(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] - b.m[i]; return r; }
__device__ __forceinline__ M3 tr(const M3 &a) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * j + i]; return r; }
__device__ __forceinline__ M3 skew(V3 q) {          // Utility::skewSymmetric, utility.hpp:36-44
    M3 r;
    r.m[0] = 0; r.m[1] = -q.z; r.m[2] = q.y; r.m[3] = q.z; r.m[4] = 0; r.m[5] = -q.x; r.m[6] = -q.y; r.m[7] = q.x; r.m[8] = 0;
    return r;
}

__device__ __forceinline__ Q4 q4(double x, double y, double z, double w) { Q4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
__device__ __forceinline__ Q4 ldq(const double *p) { return q4(p[0], p[1], p[2], p[3]); }
__device__ __forceinline__ void stq(double *p, Q4 q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {   // Hamilton product a (x) b
    return q4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
              a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
__device__ __forceinline__ Q4 qconj(Q4 a) { return q4(-a.x, -a.y, -a.z, a.w); }
__device__ __forceinline__ Q4 qinv(Q4 a) {          // Eigen inverse(): conjugate / squaredNorm
    const double n2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    return q4(-a.x / n2, -a.y / n2, -a.z / n2, a.w / n2);
}
__device__ __forceinline__ Q4 qnormalized(Q4 a) {
    const double n = sqrt(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w);
    return q4(a.x / n, a.y / n, a.z / n, a.w / n);
}
__device__ __forceinline__ M3 q2R(Q4 q) {           // Eigen toRotationMatrix()
    M3 r;
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y,
                 tzz = tz * q.z;
    r.m[0] = 1 - (tyy + tzz); r.m[1] = txy - twz; r.m[2] = txz + twy;
    r.m[3] = txy + twz; r.m[4] = 1 - (txx + tzz); r.m[5] = tyz - twx;
    r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1 - (txx + tyy);
    return r;
}
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {    // Eigen _transformVector
    const V3 u = v3(q.x, q.y, q.z);
    const V3 uv = 2.0 * cross(u, v);
    return v + q.w * uv + cross(u, uv);
}
__device__ inline Q4 R2q(const M3 &R) {             // Eigen Quaternion(Matrix3) (quaternionbase_assign_impl)
    Q4 q;
    double t = R.m[0] + R.m[4] + R.m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R.m[7] - R.m[5]) * t; q.y = (R.m[2] - R.m[6]) * t; q.z = (R.m[3] - R.m[1]) * t;
    } else {
        int i = 0;
        if (R.m[4] > R.m[0]) i = 1;
        if (R.m[8] > R.m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R.m[4 * i] - R.m[4 * j] - R.m[4 * k] + 1.0);
        double c[3];
        c[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R.m[3 * k + j] - R.m[3 * j + k]) * t;
        c[j] = (R.m[3 * j + i] + R.m[3 * i + j]) * t;
        c[k] = (R.m[3 * k + i] + R.m[3 * i + k]) * t;
        q.x = c[0]; q.y = c[1]; q.z = c[2];
    }
    return q;
}
__device__ __forceinline__ Q4 deltaQ(V3 theta) { return q4(theta.x / 2.0, theta.y / 2.0, theta.z / 2.0, 1.0); }   // utility.hpp:21-34 (first order)

// bottom-right 3x3 of Utility::Qleft(q) = w I + skew(vec);  of Utility::Qright(p) = w I - skew(vec)   (utility.hpp:56-74)
__device__ __forceinline__ M3 qleft33(Q4 q) { M3 r = skew(v3(q.x, q.y, q.z)); r.m[0] += q.w; r.m[4] += q.w; r.m[8] += q.w; return r; }
__device__ __forceinline__ M3 qright33(Q4 q) { M3 r = skew(v3(-q.x, -q.y, -q.z)); r.m[0] += q.w; r.m[4] += q.w; r.m[8] += q.w; return r; }
// bottom-right 3x3 of Qleft(a) * Qright(b):  (4x4 product) rows 1..3, cols 1..3 = vec_a * (-vec_b)^T + Ql33(a) * Qr33(b)
__device__ inline M3 qleft_qright_33(Q4 a, Q4 b) {
    M3 r = qleft33(a) * qright33(b);
    const double av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[3 * i + j] += av[i] * (-bv[j]);
    return r;
}

__device__ inline V3 R2ypr(const M3 &R) {           // utility.hpp:76-93 (degrees)
    const V3 n = v3(R.m[0], R.m[3], R.m[6]), o = v3(R.m[1], R.m[4], R.m[7]), a = v3(R.m[2], R.m[5], R.m[8]);
    const double y = atan2(n.y, n.x);
    const double p = atan2(-n.z, n.x * cos(y) + n.y * sin(y));
    const double r = atan2(a.x * sin(y) - a.y * cos(y), -o.x * sin(y) + o.y * cos(y));
    const double k = 180.0 / 3.14159265358979323846;
    return v3(y * k, p * k, r * k);
}
__device__ inline M3 ypr2R(V3 ypr) {                // utility.hpp:95-121 (degrees)
    const double k = 3.14159265358979323846 / 180.0;
    const double y = ypr.x * k, p = ypr.y * k, r = ypr.z * k;
    M3 Rz = eye3(), Ry = eye3(), Rx = eye3();
    Rz.m[0] = cos(y); Rz.m[1] = -sin(y); Rz.m[3] = sin(y); Rz.m[4] = cos(y);
    Ry.m[0] = cos(p); Ry.m[2] = sin(p); Ry.m[6] = -sin(p); Ry.m[8] = cos(p);
    Rx.m[4] = cos(r); Rx.m[5] = -sin(r); Rx.m[7] = sin(r); Rx.m[8] = cos(r);
    return Rz * Ry * Rx;
}

__device__ __forceinline__ void atomic_add(double *p, double v) { atomicAdd(p, v); }

}  // namespace be
