/* oracle/canny_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * CPU restatement of cv2.Canny(img_u8, low, high) with OpenCV's defaults
 * (apertureSize=3, L2gradient=false), the integer step of the reference's
 * forward at models/models.py:359-362.  The algorithm lives in a third-party
 * dependency of the reference (opencv-python, requirements.txt, unpinned;
 * 4.13.0 in the build image) whose source is not under /root/reference; this
 * file restates the published algorithm:
 *   1. 3x3 Sobel dx, dy (int16) with replicated borders;
 *   2. magnitude |dx|+|dy|, surrounded by a 1-pixel border of zeros;
 *   3. non-maximum suppression in 4 direction sectors decided by the fixed
 *      point tan(22.5 deg) test (TG22 = round(0.41421356 * 2^15));
 *   4. hysteresis: mag > high is an edge seed, mag > low candidates survive
 *      iff 8-connected to a seed;  output 255 / 0.
 * Pinned bit-exactly against cv2.Canny by tests/golden/make_golden.py
 * (fixtures tests/golden/canny_ref.npz) and tests/test_oracle_vs_golden.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

int canny_oracle_u8(const uint8_t* src, int rows, int cols, int low, int high, uint8_t* dst)
{
    if (rows <= 0 || cols <= 0) return -1;
    const int mstep = cols + 2;
    int16_t* dx = (int16_t*)malloc(sizeof(int16_t) * rows * cols);
    int16_t* dy = (int16_t*)malloc(sizeof(int16_t) * rows * cols);
    int* mag = (int*)calloc((size_t)(rows + 2) * mstep, sizeof(int));
    uint8_t* map = (uint8_t*)malloc((size_t)(rows + 2) * mstep);
    int* stack = (int*)malloc(sizeof(int) * (size_t)rows * cols);
    if (!dx || !dy || !mag || !map || !stack) return -2;
    memset(map, 1, (size_t)(rows + 2) * mstep);

    for (int y = 0; y < rows; ++y) {
        const uint8_t* r0 = src + (size_t)clampi(y - 1, 0, rows - 1) * cols;
        const uint8_t* r1 = src + (size_t)y * cols;
        const uint8_t* r2 = src + (size_t)clampi(y + 1, 0, rows - 1) * cols;
        for (int x = 0; x < cols; ++x) {
            int xl = clampi(x - 1, 0, cols - 1), xr = clampi(x + 1, 0, cols - 1);
            int gx = (r0[xr] + 2 * r1[xr] + r2[xr]) - (r0[xl] + 2 * r1[xl] + r2[xl]);
            int gy = (r2[xl] + 2 * r2[x] + r2[xr]) - (r0[xl] + 2 * r0[x] + r0[xr]);
            dx[y * cols + x] = (int16_t)gx;
            dy[y * cols + x] = (int16_t)gy;
            mag[(y + 1) * mstep + x + 1] = abs(gx) + abs(gy);
        }
    }

    const int TG22 = 13573;  /* (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5) */
    int sp = 0;
    for (int y = 0; y < rows; ++y) {
        const int* mp = mag + (size_t)y * mstep + 1;        /* previous row */
        const int* ma = mp + mstep;                         /* this row     */
        const int* mn = ma + mstep;                         /* next row     */
        uint8_t* pm = map + (size_t)(y + 1) * mstep + 1;
        for (int x = 0; x < cols; ++x) {
            int m = ma[x];
            int keep = 0;
            if (m > low) {
                int xs = dx[y * cols + x], ys = dy[y * cols + x];
                int ax = abs(xs), ay = abs(ys) << 15;
                int tg22x = ax * TG22;
                if (ay < tg22x) {
                    keep = (m > ma[x - 1] && m >= ma[x + 1]);
                } else {
                    int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x) {
                        keep = (m > mp[x] && m >= mn[x]);
                    } else {
                        int s = ((xs ^ ys) < 0) ? -1 : 1;
                        keep = (m > mp[x - s] && m > mn[x + s]);
                    }
                }
            }
            if (keep) {
                if (m > high) { pm[x] = 2; stack[sp++] = (y + 1) * mstep + x + 1; }
                else pm[x] = 0;
            } else pm[x] = 1;
        }
    }

    static const int dyo[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
    static const int dxo[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    while (sp > 0) {
        int p = stack[--sp];
        for (int k = 0; k < 8; ++k) {
            int q = p + dyo[k] * mstep + dxo[k];
            if (map[q] == 0) { map[q] = 2; stack[sp++] = q; }
        }
    }
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x)
            dst[(size_t)y * cols + x] = (map[(size_t)(y + 1) * mstep + x + 1] == 2) ? 255 : 0;
    free(dx); free(dy); free(mag); free(map); free(stack);
    return 0;
}
