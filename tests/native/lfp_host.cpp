// Host build of libvpx.opencl_b200/csrc/lf_packed.cuh (the plain C twins of the device
// primitives) for tests/test_lf_packed.py.  TEST INFRASTRUCTURE ONLY.
#include <stdint.h>
#include "lf_packed.cuh"

// px: n lines of 8 pixels p3 p2 p1 p0 q0 q1 q2 q3, filtered in place two lines at a time.
// kind 0 = macroblock edge, 1 = inner edge, 2 = simple (uses p1 p0 q0 q1 only), 3 = inner with
// the "no inner edges" limit (must be the identity), 4 = pack/unpack round trip.
extern "C" void lfp_run(int kind, uint8_t *px, long n, int ilim, int elim, int thr)
{
    LfPk P;
    P.ilimB = K2(ilim | 0x8000);
    P.mbEB = K2((2 * elim + 1) | 0x8000);
    P.inEB = kind == 3 ? LFP_NEVER : K2((2 * elim + 1) | 0x8000);
    P.thrB = K2(thr | 0x8000);
    for (long i = 0; i + 1 < n; i += 2) {
        uint8_t *a = px + 8 * i, *b = a + 8;
        u32 v[8];
        if (kind == 4) {
            u32 wa0, wa1, wb0, wb1, x[8];
            wa0 = a[0] | a[1] << 8 | a[2] << 16 | (u32)a[3] << 24; wa1 = a[4] | a[5] << 8 | a[6] << 16 | (u32)a[7] << 24;
            wb0 = b[0] | b[1] << 8 | b[2] << 16 | (u32)b[3] << 24; wb1 = b[4] | b[5] << 8 | b[6] << 16 | (u32)b[7] << 24;
            lfp_unpack(wa0, wb0, x[0], x[1], x[2], x[3]);
            lfp_unpack(wa1, wb1, x[4], x[5], x[6], x[7]);
            for (int k = 0; k < 8; k++) if (x[k] != ((u32)a[k] | (u32)b[k] << 16)) { a[0] ^= 0xff; return; }
            u32 ra0, rb0, ra1, rb1;
            lfp_pack(x[0], x[1], x[2], x[3], ra0, rb0);
            lfp_pack(x[4], x[5], x[6], x[7], ra1, rb1);
            if (ra0 != wa0 || rb0 != wb0 || ra1 != wa1 || rb1 != wb1) { a[0] ^= 0xff; return; }
            continue;
        }
        for (int k = 0; k < 8; k++) v[k] = (u32)a[k] | (u32)b[k] << 16;
        if (kind == 0) lfp_mbedge(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], P);
        else if (kind == 1 || kind == 3) lfp_inner(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], P);
        else lfp_simple(v[2], v[3], v[4], v[5], P.mbEB);
        for (int k = 0; k < 8; k++) { a[k] = (uint8_t)(v[k] & 0xffff); b[k] = (uint8_t)(v[k] >> 16); if ((v[k] & 0xff00ff00u)) { a[k] = b[k] = 0xAA; } }
    }
}
