// weight.cuh -- AdResS weighting functions evaluated inline on the device.
// Reference: mrmd/weighting_function/Slab.hpp:161-187, Spherical.hpp:37-75, mrmd/util/math.hpp:31-46.
#pragma once

#include "common.cuh"

namespace mrmd_b200
{
constexpr double PI = 3.14159265358979323846;

// util/math.hpp:31-46: square-and-multiply in the reference's multiplication order
__device__ __forceinline__ double powInt(double x, long long n)
{
    double ww = x;
    double yy = 1.0;
    for (long long nn = (n > 0) ? n : -n; nn != 0; nn >>= 1)
    {
        if ((nn & 1) == 1) yy *= ww;
        ww *= ww;
    }
    return (n > 0) ? yy : 1.0 / yy;
}

// Slab::operator() (Slab.hpp:161-187) / Spherical::operator() (Spherical.hpp:37-75)
__device__ __forceinline__ void weightEval(const mrmd_b200_weight& w, double x, double y, double z, double& lambda,
                                           double& modLambda, double& gx, double& gy, double& gz)
{
    gx = gy = gz = 0.0;
    if (w.kind == MRMD_B200_WEIGHT_SLAB)
    {
        const double atHalf = 0.5 * w.atRegion;
        const long long exponent = 2 * w.exponent;
        const double dx = x - w.center[0];
        const double absDx = fabs(dx);
        if (absDx < atHalf || (w.abrupt && !(absDx > atHalf + w.hyRegion)))
        {
            lambda = 1.0;
            modLambda = 1.0;
        }
        else if (absDx > atHalf + w.hyRegion)
        {
            lambda = 0.0;
            modLambda = 0.0;
        }
        else
        {
            const double arg = PI / (2.0 * w.hyRegion) * (absDx - atHalf);
            const double base = cos(arg);
            lambda = base * base;
            modLambda = powInt(base, exponent);
            const double factor =
                -PI / (2.0 * w.hyRegion) * double(exponent) * sin(arg) * powInt(base, exponent - 1) / absDx;
            gx = factor * dx;
        }
        return;
    }
    const double atRadiusSqr = w.atRegion * w.atRegion;
    const double cgRadiusSqr = (w.atRegion + w.hyRegion) * (w.atRegion + w.hyRegion);
    const double dx = x - w.center[0], dy = y - w.center[1], dz = z - w.center[2];
    const double dxSqr = dx * dx + dy * dy + dz * dz;
    if (dxSqr < atRadiusSqr)
    {
        lambda = 1.0;
        modLambda = 1.0;
        return;
    }
    if (dxSqr > cgRadiusSqr)
    {
        lambda = 0.0;
        modLambda = 0.0;
        return;
    }
    const double r = sqrt(dxSqr);
    const double arg = PI / (2.0 * w.hyRegion) * (r - w.atRegion);
    const double base = cos(arg);
    lambda = powInt(base, w.exponent);
    modLambda = lambda;
    const double factor = -PI / (2.0 * w.hyRegion) * double(w.exponent) * sin(arg) * powInt(base, w.exponent - 1) / r;
    gx = factor * dx;
    gy = factor * dy;
    gz = factor * dz;
}

// the modulated weight alone (what a pair needs from its partner): no sin, no gradient
__device__ __forceinline__ double weightModLambda(const mrmd_b200_weight& w, double x, double y, double z)
{
    if (w.kind == MRMD_B200_WEIGHT_SLAB)
    {
        const double atHalf = 0.5 * w.atRegion;
        const double absDx = fabs(x - w.center[0]);
        if (absDx < atHalf || (w.abrupt && !(absDx > atHalf + w.hyRegion))) return 1.0;
        if (absDx > atHalf + w.hyRegion) return 0.0;
        const double arg = PI / (2.0 * w.hyRegion) * (absDx - atHalf);
        return powInt(cos(arg), 2 * w.exponent);
    }
    const double atRadiusSqr = w.atRegion * w.atRegion;
    const double cgRadiusSqr = (w.atRegion + w.hyRegion) * (w.atRegion + w.hyRegion);
    const double dx = x - w.center[0], dy = y - w.center[1], dz = z - w.center[2];
    const double dxSqr = dx * dx + dy * dy + dz * dz;
    if (dxSqr < atRadiusSqr) return 1.0;
    if (dxSqr > cgRadiusSqr) return 0.0;
    const double arg = PI / (2.0 * w.hyRegion) * (sqrt(dxSqr) - w.atRegion);
    return powInt(cos(arg), w.exponent);
}
}  // namespace mrmd_b200
