// sh.cuh -- real spherical-harmonics basis (degree 0..3) and the camera-centre helper.
// Constants and coefficient order follow the reference (GR/compact.cu:554-653): sh_rest rows are
// l=1 (-y, z, -x), l=2 (xy, yz, 2zz-xx-yy, xz, xx-yy), l=3 (7 terms).
#pragma once

template <int DEG>
__device__ __forceinline__ void lgs_sh_basis(float x, float y, float z, float* b)
{
    b[0] = 0.28209479177387814f;
    if (DEG > 0) {
        const float C1 = 0.4886025119029199f;
        b[1] = -C1 * y; b[2] = C1 * z; b[3] = -C1 * x;
        if (DEG > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = 1.0925484305920792f * xy;
            b[5] = -1.0925484305920792f * yz;
            b[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
            b[7] = -1.0925484305920792f * xz;
            b[8] = 0.5462742152960396f * (xx - yy);
            if (DEG > 2) {
                b[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
                b[10] = 2.890611442640554f * xy * z;
                b[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
                b[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                b[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
                b[14] = 1.445305721320277f * z * (xx - yy);
                b[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
            }
        }
    }
}

// camera centre = -t . R^T with t = V[3,:3], R = V[:3,:3]   (GR/compact.cu:875-879)
__device__ __forceinline__ void lgs_camera_center(const float* __restrict__ Vm, float* c)
{
    float tx = -Vm[12], ty = -Vm[13], tz = -Vm[14];
    c[0] = tx * Vm[0] + ty * Vm[1] + tz * Vm[2];
    c[1] = tx * Vm[4] + ty * Vm[5] + tz * Vm[6];
    c[2] = tx * Vm[8] + ty * Vm[9] + tz * Vm[10];
}
