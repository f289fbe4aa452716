// image.cu -- image-formation kernels: echo accumulation with scatterer-volume sampling
// (main.cpp:106-144, volume.h:46-61, rfimage.h:33-40), separable PSF convolution
// (rfimage.h:93-123), envelope detection (rfimage.h:54-91) and scan conversion (rfimage.h:139,
// 183-215).  RF images live in HBM scanline-major: rf[image][element][row], one scanline contiguous.
#include "mcrt_device.cuh"
#include "mcrt_launch.h"

namespace mcrt {

namespace {

// ------------------------------------------------------------------------------------------------
// volume.h:46-61
// ------------------------------------------------------------------------------------------------
// static_cast<unsigned>(negative float) is UB in the reference; x86-64/GCC converts through a 64-bit
// integer and keeps the low 32 bits.  `% 256` of that is `& 255`.
__device__ __noinline__ uint32_t voxel_linear_exact(float x, float y, float z, float resolution)
{
    const uint32_t xi = (uint32_t)__float2ll_rz(x / resolution) & 255u;
    const uint32_t yi = (uint32_t)__float2ll_rz(y / resolution) & 255u;
    const uint32_t zi = (uint32_t)__float2ll_rz(z / resolution) & 255u;
    return (xi << 16) | (yi << 8) | zi;
}

// Same result without three IEEE divisions on the hot path: q~ = coord * (1/resolution) is within
// 2 ulp of fl(coord/resolution), so both truncate to the same integer unless q~ lies within a (much
// wider, 1e-6 relative) guard band of an integer -- only then are the exact quotients evaluated.
// `risky` is OR-ed with "the fast result may be wrong"; the caller then re-evaluates with voxel_linear_exact.
__device__ __forceinline__ uint32_t voxel_linear_fast(float3 p, float inv_resolution, int& risky)
{
    const float qx = p.x * inv_resolution, qy = p.y * inv_resolution, qz = p.z * inv_resolution;
    const int ix = __float2int_rz(qx), iy = __float2int_rz(qy), iz = __float2int_rz(qz);
    const float ax = fabsf(qx - (float)ix), ay = fabsf(qy - (float)iy), az = fabsf(qz - (float)iz);   // [0,1) unless saturated
    const float amin = fminf(fminf(ax, ay), az), amax = fmaxf(fmaxf(ax, ay), az);
    const float qmax = fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz));
    const float eps = qmax * 1e-6f + 1e-30f;
    // branch-free: the three conditions are evaluated as predicates and OR-ed into the flag
    risky |= (int)(!(amin >= eps)) | (int)(!(amax <= 1.0f - eps)) | (int)(!(qmax < 2.0e9f));
    return ((uint32_t)(ix & 255) << 16) | ((uint32_t)(iy & 255) << 8) | (uint32_t)(iz & 255);
}

__device__ __forceinline__ uint32_t voxel_linear(float3 p, float resolution, float inv_resolution)
{
    const float qx = p.x * inv_resolution, qy = p.y * inv_resolution, qz = p.z * inv_resolution;
    const int ix = __float2int_rz(qx), iy = __float2int_rz(qy), iz = __float2int_rz(qz);
    const float ax = fabsf(qx - (float)ix), ay = fabsf(qy - (float)iy), az = fabsf(qz - (float)iz);   // [0,1) unless saturated
    const float amin = fminf(fminf(ax, ay), az), amax = fmaxf(fmaxf(ax, ay), az);
    const float qmax = fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz));
    const float eps = qmax * 1e-6f + 1e-30f;
    if (!(amin >= eps && amax <= 1.0f - eps && qmax < 2.0e9f)) return voxel_linear_exact(p.x, p.y, p.z, resolution);
    return ((uint32_t)(ix & 255) << 16) | ((uint32_t)(iy & 255) << 8) | (uint32_t)(iz & 255);
}

// Correctly rounded coord / resolution from three FP32 instructions (Markstein's FMA division step):
//   q0 = coord * y (y = fl(1/resolution)),  r = fma(-q0, resolution, coord) (exact residual),  q = fma(r, y, q0).
// For a fixed divisor this equals the IEEE quotient except where the residual underflows (|coord| < ~1e-31), and
// there both truncate to 0.  It is not taken on faith: k_validate_fma_division checks, for the context's actual
// resolution and ALL 2^32 float bit patterns, that the voxel index byte equals the one from the IEEE division
// (a few ms at mcrt_create); a resolution that fails keeps the guarded path above.
__device__ __forceinline__ float div_fma(float x, float resolution, float y)
{
    const float q0 = x * y;
    const float r = __fmaf_rn(-q0, resolution, x);
    return __fmaf_rn(r, y, q0);
}

// |q| < 2^31 on all three axes is established per segment by the caller, so the 32-bit conversion is the
// truncation of volume.h:49-51 and `& 255` its `% 256`.
__device__ __forceinline__ uint32_t voxel_linear_fma(float3 p, float resolution, float y)
{
    const int ix = __float2int_rz(div_fma(p.x, resolution, y));
    const int iy = __float2int_rz(div_fma(p.y, resolution, y));
    const int iz = __float2int_rz(div_fma(p.z, resolution, y));
    // (ix & 255) << 16 | (iy & 255) << 8 | (iz & 255) as two byte permutes
    const uint32_t zy = __byte_perm((uint32_t)iz, (uint32_t)iy, 0x0040);          // b0 = iz.b0, b1 = iy.b0 (b2, b3 dropped below)
    return __byte_perm(zy, (uint32_t)ix & 255u, 0x7410);                          // b2 = ix.b0, b3 = 0
}

__global__ void __launch_bounds__(256) k_validate_fma_division(const float resolution, unsigned int* __restrict__ mismatches)
{
    const float y = 1.0f / resolution;
    unsigned int bad = 0;
    // thread t checks the 256 bit patterns t * 256 .. t * 256 + 255
    const uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) << 8;
    for (uint32_t k = 0; k < 256u; k++) {
        const float x = __uint_as_float(base + k);
        const float qe = x / resolution;
        if (!(fabsf(qe) < 2147483648.0f)) continue;      // beyond 2^31 (or non-finite) the caller never takes the FMA path
        const float q = div_fma(x, resolution, y);
        bad += (__float2int_rz(qe) & 255) != (__float2int_rz(q) & 255);
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ------------------------------------------------------------------------------------------------
// Echo accumulation (main.cpp:106-144 + rf_image::add_echo, rfimage.h:33-40), no atomics.
//
// k_accumulate: one thread marches one Monte-Carlo path -- its segments in order, each a sequential
// fp32 position chain / fp64 time chain that is reproduced exactly.  The RF row a step lands in
// grows (almost always) by one per step, so the thread keeps the running sum of the current row in
// a register and writes every row of its private column exactly once, in order, to HBM:
// columns[scanline][row][sample] (samples of a scanline adjacent -> the 16 lanes that march a
// scanline in lock-step store one 64-byte line).  A row that is revisited (time going backwards,
// possible only with spacing != 1) falls back to a read-modify-write of the thread's own column.
// Scatterer-volume gathers are issued MCRT_ACC_UNROLL steps ahead of their use.
//
// k_reduce_samples: rf[scanline][row] = sum over samples in sample order (fixed order -> results
// do not depend on scheduling; N-GPU sharding is bit-identical to 1 GPU).
// ------------------------------------------------------------------------------------------------
#ifndef MCRT_ACC_UNROLL
#define MCRT_ACC_UNROLL 6      // measured: 4 -> 1.51 ms, 6 -> 1.42, 8 -> 1.42 ms per 256 frames (64 registers each)
#endif

// Slow path of the column writer: the row being closed is not the next unwritten one.  Returns the
// new `written` watermark.  Kept out of line so the hot loop stays small.
__device__ __noinline__ int column_flush_slow(float* col, int S, int written, int cur_row, float cur_acc)
{
    if (cur_row < 0) return written;
    if (cur_row >= written) {
        for (int r = written; r < cur_row; r++) col[(size_t)r * S] = 0.0f;              // rows nothing landed in
        col[(size_t)cur_row * S] = cur_acc;
        return cur_row + 1;
    }
    col[(size_t)cur_row * S] += cur_acc;                                                 // revisited row (own data)
    return written;
}

// Per-thread writer of one private RF column; all members live in registers (everything inlines).
struct ColumnWriter {
    float* col;          // row r lives at col[r * S]
    float* wptr;         // = col + written * S: where the next in-order row goes
    int S, rows;
    int written;         // rows [0, written) have been stored
    int cur_row;         // row being accumulated in cur_acc (-1: none yet)
    float cur_acc;
    double row_period, inv_row_period;

    // rf_image::add_echo (rfimage.h:33-40): row = micros / (axial_resolution_/speed_of_sound_), truncated;
    // dropped when row >= max_rows.  The product with the reciprocal decides the row unless it lands
    // within 1e-9 of an integer, where the exact IEEE quotient is taken instead.
    __device__ __forceinline__ void add_echo(float echo, double micros)
    {
        double rowd = micros * inv_row_period;
        int row = __double2int_rd(rowd);
        const double fr = rowd - (double)row;
        if (!(fr >= 1e-9 && fr <= 1.0 - 1e-9)) {
            rowd = micros / row_period;
            if (!(rowd < (double)(unsigned)rows)) return;
            row = (int)rowd;
        } else if (row >= rows) {
            return;
        }
        add_row(echo, row);
    }
    // the echo lands in RF row `row` (0 <= row < rows already established)
    __device__ __forceinline__ void add_row(float echo, int row)
    {
        if (row == cur_row) { cur_acc += echo; return; }
        if (cur_row == written) { *wptr = cur_acc; wptr += S; written++; }                        // the common case: next row
        else { written = column_flush_slow(col, S, written, cur_row, cur_acc); wptr = col + (size_t)written * S; }
        cur_row = row;
        cur_acc = echo;                                                                   // 0 + echo
    }
    __device__ __forceinline__ void finish()
    {
        written = column_flush_slow(col, S, written, cur_row, cur_acc);
        for (int r = written; r < rows; r++) col[(size_t)r * S] = 0.0f;                  // rf_image.clear(), main.cpp:102
    }
};

#ifndef MCRT_ACC_MIN_CTAS
#define MCRT_ACC_MIN_CTAS 8      // 64 registers, 32 warps/SM: measured 9 % faster than 6 (78 registers)
#endif
template <bool FMADIV>
__global__ void __launch_bounds__(128, MCRT_ACC_MIN_CTAS) k_accumulate(const SceneDev sc, const AcqDev aq, const float2* __restrict__ volume,
                                                   const DevSegment* __restrict__ segments, const int32_t* __restrict__ nseg,
                                                   const int n_paths, float* __restrict__ columns,
                                                   unsigned long long* __restrict__ steps_total)
{
    __shared__ DevMaterial s_mat[MCRT_MAX_SMEM_MATERIALS];
    for (int i = threadIdx.x; i < sc.n_mat && i < MCRT_MAX_SMEM_MATERIALS; i += blockDim.x) s_mat[i] = sc.materials[i];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long my_steps = 0;
    if (p < n_paths) {
        const int S = aq.samples;
        const int scanline = p / S;
        ColumnWriter w;
        w.S = S; w.rows = aq.rows;
        w.col = columns + (size_t)scanline * aq.rows * S + (p - scanline * S);
        w.wptr = w.col; w.written = 0; w.cur_row = -1; w.cur_acc = 0.0f;
        w.row_period = aq.row_period_us; w.inv_row_period = aq.inv_row_period;
        const float axres_f = aq.axres_f;
        const double time_step = aq.time_step_us;
        const double inv_time_step = 1.0 / aq.time_step_us;
        const double max_travel_time = aq.max_travel_time_us;
        const float samples_f = (float)(size_t)S;
        const float vres = aq.vol_resolution, inv_vres = 1.0f / aq.vol_resolution;
        // One march step advances the RF row by `ratio` = time_step / row_period (1.00069 for the reference's
        // 322.22 um step on a 322 um row grid).  When 1 <= ratio < 1.2 the rows of an unrolled block are
        // row0, row0+1, ... as long as the fractional row position of its first step stays below
        // block_safe_hi; then no per-step fp64 row computation is needed (guard 1e-6 >> the ~1e-12 rounding
        // of the iterated fp64 time chain and of the product with the reciprocal).
        const double row_delta = aq.time_step_us * aq.inv_row_period - 1.0;
        const bool fast_rows = row_delta >= 0.0 && row_delta < 0.2;
        const double block_safe_hi = 1.0 - 1e-6 - (double)(MCRT_ACC_UNROLL - 1) * row_delta;

        const int ns = nseg[p];
        for (int k = 0; k < ns; k++) {
            const DevSegment* sg = segments + (size_t)p * aq.max_depth + k;
            const float4 s0 = __ldg(&sg->s0), s1 = __ldg(&sg->s1), s2 = __ldg(&sg->s2);
            const int4 s3 = __ldg(&sg->s3);
            const DevMaterial media = s_mat[s3.z];
            const double distance_traveled = __hiloint2double(s3.y, s3.x);
            const double starting_micros = ((distance_traveled * 1000) / 1) / aq.speed;      // main.cpp:114
            const float3 from = make_float3(s0.x, s0.y, s0.z), to = make_float3(s2.x, s2.y, s2.z);
            const double distance = (double)(v_length(v_sub(to, from)) * 10.0f);               // scene.cpp:342-346
            const double steps_d = distance / aq.axres_mm;                                      // main.cpp:116
            unsigned long long steps64;                                                         // B-14
            if (!(steps_d >= 0.0)) steps64 = 0;
            else if (steps_d >= 9.0e18) steps64 = 9000000000000000000ULL;
            else steps64 = (unsigned long long)steps_d;
            const uint32_t steps32 = (uint32_t)steps64;
            const float3 delta_step = v_scl(make_float3(s1.x, s1.y, s1.z), axres_f);            // main.cpp:117
            float3 point = from;
            double time_elapsed = starting_micros;
            float intensity = s1.w;
            const float decay = mc_expf(-s2.w * axres_f * 0.01f * aq.frequency * 1.0f);         // main.cpp:135
            // The loop of main.cpp:124 runs while (step < steps && time_elapsed < max_travel_time); the time
            // bound caps it far below 2^31 iterations, so the step budget fits an int.
            int remaining = steps64 > 0x7fffffffULL ? 0x7fffffff : (int)steps64;
            // steps that certainly satisfy the time bound (2 steps of slack against the rounding of the
            // iterated fp64 sum) run in unrolled blocks without per-step checks
            const double safe_d = (max_travel_time - time_elapsed) * inv_time_step - 2.0;
            int n_safe = safe_d > 0.0 ? (safe_d < 2.0e9 ? (int)safe_d : 2000000000) : 0;
            if (n_safe > remaining) n_safe = remaining;
            // a medium with sigma == 0 and mu0 == 0 (coupling gel) scatters exactly +0 at every voxel: adding +0
            // never changes a row, so only the time chain (which bounds the step count) is advanced
            if (media.sigma == 0.0f && media.mu0 == 0.0f) {
                while (remaining > 0 && time_elapsed < max_travel_time) { time_elapsed = time_elapsed + time_step; remaining--; my_steps++; }
                n_safe = 0;
            }
            const int n_blocks = n_safe / MCRT_ACC_UNROLL;
            // FMA-division voxel indices need |coord / resolution| < 2^31 along the whole march: the march is a
            // straight line of n_safe steps of length <= axres per axis from `from` (bound with 2x slack)
            const float reach = fmaxf(fmaxf(fabsf(from.x), fabsf(from.y)), fabsf(from.z)) + 2.0f * (float)n_safe * fabsf(axres_f) *
                                fmaxf(fmaxf(fabsf(s1.x), fabsf(s1.y)), fabsf(s1.z));
            const bool fma_ok = FMADIV && reach * inv_vres < 1.0e9f;
            for (int b = 0; b < n_blocks; b++) {
                uint32_t idx[MCRT_ACC_UNROLL];
                if (fma_ok) {
#pragma unroll
                    for (int u = 0; u < MCRT_ACC_UNROLL; u++) {
                        idx[u] = voxel_linear_fma(point, vres, inv_vres);
                        point = v_add(point, delta_step);                                       // main.cpp:131
                    }
                } else {
                    float3 pts[MCRT_ACC_UNROLL];
                    int risky = 0;                      // one guard branch per block instead of one per step
#pragma unroll
                    for (int u = 0; u < MCRT_ACC_UNROLL; u++) {
                        pts[u] = point;
                        idx[u] = voxel_linear_fast(point, inv_vres, risky);
                        point = v_add(point, delta_step);                                       // main.cpp:131
                    }
                    if (risky) {
#pragma unroll
                        for (int u = 0; u < MCRT_ACC_UNROLL; u++) idx[u] = voxel_linear_exact(pts[u].x, pts[u].y, pts[u].z, vres);
                    }
                }
                float2 vox[MCRT_ACC_UNROLL];
#pragma unroll
                for (int u = 0; u < MCRT_ACC_UNROLL; u++) vox[u] = __ldg(&volume[idx[u]]);      // (noise, probability)
                const double rowd0 = time_elapsed * w.inv_row_period;
                const int row0 = __double2int_rd(rowd0);
                const double f0 = rowd0 - (double)row0;
                if (fast_rows && f0 >= 1e-6 && f0 <= block_safe_hi && row0 + MCRT_ACC_UNROLL <= w.rows) {
                    float echo[MCRT_ACC_UNROLL];
#pragma unroll
                    for (int u = 0; u < MCRT_ACC_UNROLL; u++) {
                        // get_scattering(mu1, mu0, sigma, ...): density = mu1, mu = mu0 (main.cpp:126 vs volume.h:46)
                        const float scattering = vox[u].y >= media.mu1 ? vox[u].x * media.sigma + media.mu0 : 0.0f;
                        echo[u] = intensity * scattering;
                        time_elapsed = time_elapsed + time_step;
                        intensity *= decay;
                    }
                    w.add_row(echo[0], row0);
                    if (w.written == row0) {
                        // in-order streaming: rows row0 .. row0+U-1 receive exactly one echo each from this
                        // block; every closed row goes straight to the column, the last one stays in the register
#pragma unroll
                        for (int u = 1; u < MCRT_ACC_UNROLL; u++) {
                            *w.wptr = w.cur_acc;
                            w.wptr += w.S;
                            w.cur_acc = echo[u];
                        }
                        w.written = row0 + MCRT_ACC_UNROLL - 1;
                        w.cur_row = row0 + MCRT_ACC_UNROLL - 1;
                    } else {
#pragma unroll
                        for (int u = 1; u < MCRT_ACC_UNROLL; u++) w.add_row(echo[u], row0 + u);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < MCRT_ACC_UNROLL; u++) {
                        const float scattering = vox[u].y >= media.mu1 ? vox[u].x * media.sigma + media.mu0 : 0.0f;
                        w.add_echo(intensity * scattering, time_elapsed);
                        time_elapsed = time_elapsed + time_step;
                        intensity *= decay;
                    }
                }
            }
            remaining -= n_blocks * MCRT_ACC_UNROLL;
            my_steps += (unsigned long long)(n_blocks * MCRT_ACC_UNROLL);
            while (remaining > 0 && time_elapsed < max_travel_time) {                           // checked tail
                const float2 vox = __ldg(&volume[voxel_linear(point, vres, inv_vres)]);
                const float scattering = vox.y >= media.mu1 ? vox.x * media.sigma + media.mu0 : 0.0f;
                w.add_echo(intensity * scattering, time_elapsed);
                point = v_add(point, delta_step);
                time_elapsed = time_elapsed + time_step;
                intensity *= decay;
                remaining--;
                my_steps++;
            }
            // main.cpp:139
            w.add_echo(s0.w / samples_f, starting_micros + time_step * (double)(uint32_t)(steps32 - 1u));
        }
        w.finish();
    }
    if (steps_total) {
        for (int off = 16; off > 0; off >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, off);
        if ((threadIdx.x & 31) == 0 && my_steps) atomicAdd(steps_total, my_steps);
    }
}

// ------------------------------------------------------------------------------------------------
// Windowed echo accumulation: the same march as k_accumulate, but the private columns never reach HBM.
//
// A *group* of threads owns G whole scanlines (thread = one Monte-Carlo path, as before): a warp when S <= 32
// (G = 32 / S; the group then synchronises with __syncwarp, which costs nothing -- the CTA-wide barrier of the first
// version was its largest stall), the whole CTA for 32 < S <= 128 (G = 128 / S).  Time along a path is monotone
// (spacing == 1: every segment starts where the previous one ended), so all paths of the group sweep the RF rows
// together: the rows are processed in windows of MCRT_WIN_ROWS; inside a window every thread runs its march until
// its next echo would land beyond the window, writing its private column of the window to shared memory (each row
// once, in order: plain stores, no read-modify-write, no atomics).  Then the group sums the S columns of each
// scanline in sample order -- the exact order of k_reduce_samples, so both paths give bit-identical images -- and
// streams the finished rows to rf[scanline][row] with coalesced stores.  Compared with k_accumulate +
// k_reduce_samples this removes the write and the re-read of the [scanline][row][sample] columns (2 x 1.95 GB per
// 256 frames of 256 x 16 paths) and one launch.
// The host only selects this kernel for spacing == (1,1,1) and S <= 128; an echo that nevertheless arrives for an
// already finished window is added with atomicAdd and counted in `late_echoes` (tests hold the counter at 0).
// ------------------------------------------------------------------------------------------------
#ifndef MCRT_WIN_RING
#define MCRT_WIN_RING 32           // rows of the shared-memory ring (power of two)
#endif
#ifndef MCRT_WIN_UNROLL
#define MCRT_WIN_UNROLL 8          // steps per unrolled block
#endif
// A window finishes MCRT_WIN_ROWS rows; an unrolled block only has to START inside the window, its last rows may run
// up to MCRT_WIN_UNROLL - 1 rows ahead into the ring's slack (they belong to the next window and stay in the ring).
#ifndef MCRT_WIN_EDGE_SKIP
#define MCRT_WIN_EDGE_SKIP 1       // a lane whose next unrolled block starts beyond the window waits for the next window at once (see the checked-step loop)
#endif
#ifndef MCRT_WIN_PIN_COL
#define MCRT_WIN_PIN_COL 0         // 1: the lane's ring-column address is pinned in a register (compile-time sample count only)
#endif
#ifndef MCRT_WIN_TAIL_BLOCK
#define MCRT_WIN_TAIL_BLOCK 2      // > 0: segment tails of at least this many steps are taken as one masked unrolled block (see the kernel)
#endif
#define MCRT_WIN_ROWS (MCRT_WIN_RING - MCRT_WIN_UNROLL)
#define MCRT_WIN_SLOT(row) ((row) & (MCRT_WIN_RING - 1))

// Cold paths of the windowed writer, out of line so the hot loop stays small (instruction-cache footprint matters here:
// profiles/r01_traversal_ab.txt).  An echo for a row of an already finished window (time ran backwards):
__device__ __noinline__ void win_late_echo(float* rf_px, float echo, unsigned long long* late_echoes)
{
    atomicAdd(rf_px, echo);
    atomicAdd(late_echoes, 1ULL);
}
// ... and for an already closed row of the thread's own column in the current window; returns the new `written`
__device__ __noinline__ int win_revisit_row(float* my_col, int stride, int base, int written, int cur_row, int row, float echo)
{
    for (int r = written > base ? written : base; r < cur_row; r++) my_col[((r) & (MCRT_WIN_RING - 1)) * stride] = 0.0f;
    my_col[((row) & (MCRT_WIN_RING - 1)) * stride] += echo;
    return cur_row;
}

// ring row stride (floats) of a group with T active threads: a multiple of 4 (float4 reduce loads) whose quarter is odd
// (conflict-free quarter-warp phases for consecutive rows)
__host__ __device__ __forceinline__ int win_stride(int T) { const int q = (T + 3) / 4; return 4 * (q | 1); }

// SCT > 0 (WARP only): the number of samples per element as a compile-time constant dividing 32 (16 in BASELINE's configuration):
// the ring stride, the ring stores of an unrolled block and the sample reduction then carry no run-time index arithmetic.
// TREE (ray-tree mode, WARP and SCT = 32): the warp owns ONE scanline and a lane marches ONE SEGMENT of its ray trees per round (a
// segment is self-contained: start point, start time, initial intensity), 32 segments per round in (level, path, node) order
// (tree_first / tree_end: the scanline's segment range per level, TreeBuffers).  The 32 columns of a round are summed in lane order
// and ADDED to the scanline's RF rows (rf is cleared before the launch); windows no lane of the round touches are skipped.
template <bool FMADIV, bool WARP, int SCT, bool TREE = false>
__global__ void __launch_bounds__(128, MCRT_ACC_MIN_CTAS) k_accumulate_win(const SceneDev sc, const AcqDev aq, const float2* __restrict__ volume,
                                                       const DevSegment* __restrict__ segments, const int32_t* __restrict__ nseg,
                                                       const int n_scanlines, const int G_rt, float* __restrict__ rf,
                                                       unsigned long long* __restrict__ steps_total,
                                                       unsigned long long* __restrict__ late_echoes,
                                                       const int* __restrict__ tree_first = nullptr, const int* __restrict__ tree_end = nullptr,
                                                       const int tree_levels = 0, const int tree_stride = 0)
{
    extern __shared__ float4 s_win4[];                     // per group: [MCRT_WIN_RING][stride]; row r lives in slot r % RING
    __shared__ DevMaterial s_mat[MCRT_MAX_SMEM_MATERIALS];
    for (int i = threadIdx.x; i < sc.n_mat && i < MCRT_MAX_SMEM_MATERIALS; i += blockDim.x) s_mat[i] = sc.materials[i];
    const int S = SCT > 0 ? SCT : aq.samples, rows = aq.rows;
    const int G = SCT > 0 ? 32 / SCT : G_rt;
    const int rf_pitch = aq.rf_pitch;                      // row stride of rf (>= rows; padded to 16 B for the TMA-staged post kernel)
    const int T = G * S;                                   // active threads of a group
    const int stride = SCT > 0 ? 36 /* = win_stride(32) */ : win_stride(T);
    const int group = WARP ? (int)(blockIdx.x * 4 + (threadIdx.x >> 5)) : (int)blockIdx.x;
    const int t = WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x;          // index within the group
    float* const s_win = reinterpret_cast<float*>(s_win4) + (WARP ? (size_t)(threadIdx.x >> 5) * MCRT_WIN_RING * stride : 0);
    auto group_sync = [&]() { if (WARP) __syncwarp(); else __syncthreads(); };
    const int scanline0 = group * G;
    const int my_scanline = scanline0 + t / S;
    const bool active = t < T && my_scanline < n_scanlines;
    const int p = my_scanline * S + (t - (t / S) * S);
    __syncthreads();

    const float axres_f = aq.axres_f;
    const double time_step = aq.time_step_us;
    const double inv_time_step = 1.0 / aq.time_step_us;
    const double max_travel_time = aq.max_travel_time_us;
    const double row_period = aq.row_period_us, inv_row_period = aq.inv_row_period;
    const float samples_f = (float)(size_t)(TREE ? aq.samples : S);   // (TREE: the 32 lanes of the group are segments, not samples)
    const float vres = aq.vol_resolution, inv_vres = 1.0f / aq.vol_resolution;
    const double row_delta = aq.time_step_us * aq.inv_row_period - 1.0;
    const bool fast_rows = row_delta >= 0.0 && row_delta < 0.2;
    const double block_safe_hi = 1.0 - 1e-6 - (double)(MCRT_WIN_UNROLL - 1) * row_delta;

    // TREE: the scanline's segments, level by level; round r gives lane t the (32 r + t)-th of them
    int tree_total = 0;
    if (TREE && scanline0 < n_scanlines)
        for (int l = 0; l < tree_levels; l++) { const int e = tree_end[(size_t)l * tree_stride + scanline0]; if (e) tree_total += e - tree_first[(size_t)l * tree_stride + scanline0]; }
    unsigned long long my_steps = 0;
    for (int round0 = 0; round0 < (TREE ? tree_total : 1); round0 += 32) {
    int tree_slot = -1;
    if (TREE) {
        int j = round0 + t;
        for (int l = 0; l < tree_levels && tree_slot < 0; l++) {
            const int e = tree_end[(size_t)l * tree_stride + scanline0];
            const int f = e ? tree_first[(size_t)l * tree_stride + scanline0] : 0;
            if (j < e - f) tree_slot = f + j; else j -= e - f;
        }
    }
    // ---- per-path march state, kept in registers across windows ----
    const int ns = TREE ? (tree_slot >= 0 ? 1 : 0) : (active ? nseg[p] : 0);
    int pend_row = -1;               // TREE: the RF row of the lane's next echo once it is known to lie beyond the current window
    if (TREE && tree_slot >= 0) {
        // a lower bound of the segment's first echo row (the exact row is found by try_echo): lets the warp skip the windows before it
        const int4 s3 = __ldg(&segments[tree_slot].s3);
        const double row_lo = ((__hiloint2double(s3.y, s3.x) * 1000) / aq.speed) * inv_row_period - 2.0;
        pend_row = row_lo > 0.0 ? (row_lo < 2.0e9 ? (int)row_lo : 2000000000) : 0;
    }
    int k = 0;                       // next segment to load
    bool in_seg = false;             // a segment is loaded and not finished
    bool fma_ok = false;
    float3 point = make_float3(0.f, 0.f, 0.f), delta_step = point;
    double time_elapsed = 0.0;
    float intensity = 0.0f, decay = 0.0f;
    float m_mu0 = 0.0f, m_mu1 = 0.0f, m_sigma = 0.0f;
    int remaining = 0;               // steps left by the step budget (main.cpp:124 `step < steps`)
    int n_safe = 0;                  // of those, steps that certainly satisfy the time bound
    float end_echo = 0.0f;           // the segment's closing echo (main.cpp:139) ...
    double end_micros = 0.0;         // ... and its time
    // column writer state: rows [.., written) of my column are stored; cur_row accumulates in cur_acc
    int written = 0, cur_row = -1;
    float cur_acc = 0.0f;
    float* const my_col = s_win + t;
#if MCRT_WIN_PIN_COL
    // the shared-memory address of the lane's column, held in a register: left to itself the compiler re-derives it from the thread index
    // (nine instructions) in front of the ring stores of every unrolled block
    unsigned col_addr = (unsigned)__cvta_generic_to_shared(my_col);
    asm volatile("" : "+r"(col_addr));
#endif

    for (int base = 0; base < rows; base += MCRT_WIN_ROWS) {
        const int wend = base + MCRT_WIN_ROWS < rows ? base + MCRT_WIN_ROWS : rows;
        if (TREE) {
            // nothing of this round lands in the window: no lane has a pending row in it, and every lane is finished or knows that
            // its next echo lies beyond it
            const bool idle = cur_row < 0 && ((k >= ns && !in_seg) || pend_row >= wend);
            if (__all_sync(0xffffffffu, idle)) { if (written < wend) written = wend; continue; }
        }
        // put `echo` into RF row `row` of my column (base <= row < wend, row >= cur_row)
        auto add_row = [&](float echo, int row) {
            if (row == cur_row) { cur_acc += echo; return; }
            if (cur_row >= base) {
                for (int r = written > base ? written : base; r < cur_row; r++) my_col[MCRT_WIN_SLOT(r) * stride] = 0.0f;   // rows nothing landed in
                my_col[MCRT_WIN_SLOT(cur_row) * stride] = cur_acc;
                written = cur_row + 1;
            }
            cur_row = row;
            cur_acc = echo;                                                                     // 0 + echo
        };
        // rf_image::add_echo (rfimage.h:33-40) with the exact-row guard of ColumnWriter::add_echo.  Returns false when
        // the echo belongs to a later window (nothing consumed).
        auto try_echo = [&](float echo, double micros) -> bool {
            double rowd = micros * inv_row_period;
            int row = __double2int_rd(rowd);
            const double fr = rowd - (double)row;
            if (!(fr >= 1e-9 && fr <= 1.0 - 1e-9)) {
                rowd = micros / row_period;
                if (!(rowd < (double)(unsigned)rows)) return true;                              // dropped (row >= max_rows)
                row = (int)rowd;
            } else if (row >= rows) {
                return true;
            }
            if (row >= wend) { if (TREE) pend_row = row; return false; }
            if (row < base) {                                                                   // a finished window: see header
                win_late_echo(&rf[(size_t)my_scanline * rf_pitch + row], echo, late_echoes);
                return true;
            }
            if (row < cur_row) {                                                                // revisited row of my own column
                written = win_revisit_row(my_col, stride, base, written, cur_row, row, echo);
                return true;
            }
            add_row(echo, row);
            return true;
        };

        bool waiting = !active || (TREE && pend_row >= wend);   // true: my next echo lies beyond this window (or the path is done)
        while (!waiting) {
            if (!in_seg) {
                if (k >= ns) break;
                const DevSegment* sg = TREE ? segments + tree_slot : segments + (size_t)p * aq.max_depth + k;
                const float4 s0 = __ldg(&sg->s0), s1 = __ldg(&sg->s1), s2 = __ldg(&sg->s2);
                const int4 s3 = __ldg(&sg->s3);
                const DevMaterial media = s_mat[s3.z];
                m_mu0 = media.mu0; m_mu1 = media.mu1; m_sigma = media.sigma;
                const double distance_traveled = __hiloint2double(s3.y, s3.x);
                const double starting_micros = ((distance_traveled * 1000) / 1) / aq.speed;      // main.cpp:114
                const float3 from = make_float3(s0.x, s0.y, s0.z), to = make_float3(s2.x, s2.y, s2.z);
                const double distance = (double)(v_length(v_sub(to, from)) * 10.0f);               // scene.cpp:342-346
                const double steps_d = distance / aq.axres_mm;                                      // main.cpp:116
                unsigned long long steps64;                                                         // B-14
                if (!(steps_d >= 0.0)) steps64 = 0;
                else if (steps_d >= 9.0e18) steps64 = 9000000000000000000ULL;
                else steps64 = (unsigned long long)steps_d;
                const uint32_t steps32 = (uint32_t)steps64;
                delta_step = v_scl(make_float3(s1.x, s1.y, s1.z), axres_f);                      // main.cpp:117
                point = from;
                time_elapsed = starting_micros;
                intensity = s1.w;
                decay = mc_expf(-s2.w * axres_f * 0.01f * aq.frequency * 1.0f);                  // main.cpp:135
                remaining = steps64 > 0x7fffffffULL ? 0x7fffffff : (int)steps64;
                const double safe_d = (max_travel_time - time_elapsed) * inv_time_step - 2.0;
                n_safe = safe_d > 0.0 ? (safe_d < 2.0e9 ? (int)safe_d : 2000000000) : 0;
                if (n_safe > remaining) n_safe = remaining;
                end_echo = s0.w / samples_f;
                end_micros = starting_micros + time_step * (double)(uint32_t)(steps32 - 1u);
                const float reach = fmaxf(fmaxf(fabsf(from.x), fabsf(from.y)), fabsf(from.z)) +
                                    2.0f * (float)n_safe * fabsf(axres_f) * fmaxf(fmaxf(fabsf(s1.x), fabsf(s1.y)), fabsf(s1.z));
                fma_ok = FMADIV && reach * inv_vres < 1.0e9f;
                // coupling gel (sigma == mu0 == 0) scatters exactly +0 everywhere: only the time chain, which bounds the
                // step count, is advanced -- in one go, it touches no row
                if (m_sigma == 0.0f && m_mu0 == 0.0f) {
                    while (remaining > 0 && time_elapsed < max_travel_time) { time_elapsed = time_elapsed + time_step; remaining--; my_steps++; }
                    remaining = 0; n_safe = 0;
                }
                in_seg = true;
                k++;
            }
            // ---- unrolled blocks: MCRT_WIN_UNROLL steps whose rows are row0 .. row0+U-1, all inside this window ----
            while (n_safe >= MCRT_WIN_UNROLL) {
                const double rowd0 = time_elapsed * inv_row_period;
                const int row0 = __double2int_rd(rowd0);
                const double f0 = rowd0 - (double)row0;
                if (!(fast_rows && f0 >= 1e-6 && f0 <= block_safe_hi && row0 < wend && row0 + MCRT_WIN_UNROLL <= rows && row0 >= base && row0 >= cur_row)) break;
                uint32_t idx[MCRT_WIN_UNROLL];
                if (FMADIV) {
                    // a segment that could leave the range of the 32-bit conversion (never in practice) takes the checked
                    // single-step path instead: keeps the guarded-reciprocal code out of this kernel's hot loop
                    if (!fma_ok) break;
#pragma unroll
                    for (int u = 0; u < MCRT_WIN_UNROLL; u++) { idx[u] = voxel_linear_fma(point, vres, inv_vres); point = v_add(point, delta_step); }
                } else {
#pragma unroll
                    for (int u = 0; u < MCRT_WIN_UNROLL; u++) { idx[u] = voxel_linear(point, vres, inv_vres); point = v_add(point, delta_step); }
                }
                float2 vox[MCRT_WIN_UNROLL];
#pragma unroll
                for (int u = 0; u < MCRT_WIN_UNROLL; u++) vox[u] = __ldg(&volume[idx[u]]);      // (noise, probability)
                float echo[MCRT_WIN_UNROLL];
#pragma unroll
                for (int u = 0; u < MCRT_WIN_UNROLL; u++) {
                    // get_scattering(mu1, mu0, sigma, ...): density = mu1, mu = mu0 (main.cpp:126 vs volume.h:46)
                    const float scattering = vox[u].y >= m_mu1 ? vox[u].x * m_sigma + m_mu0 : 0.0f;
                    echo[u] = intensity * scattering;
                    time_elapsed = time_elapsed + time_step;
                    intensity *= decay;
                }
                const int slot0 = MCRT_WIN_SLOT(row0);
                // the two common cases, neither wrapping in the ring: the block continues the column row by row (a pending row right
                // below row0), or it opens the window (nothing pending, the column is written up to row0) -- the pending row and the
                // first U - 1 echoes go to consecutive ring rows at constant offsets, the last echo stays pending
                const bool cont = cur_row + 1 == row0 && written == cur_row && cur_row >= base && slot0 >= 1;
                const bool fresh = cur_row < 0 && written == row0;
                if ((cont || fresh) && slot0 + MCRT_WIN_UNROLL <= MCRT_WIN_RING) {
#if MCRT_WIN_PIN_COL
                    if (SCT > 0) {
                        const unsigned dsta = col_addr + (unsigned)slot0 * (unsigned)(36 * 4);
                        if (cont) asm volatile("st.shared.f32 [%0+-144], %1;" :: "r"(dsta), "f"(cur_acc) : "memory");
                        static_assert(MCRT_WIN_UNROLL == 8, "seven pinned ring stores");
#define MCRT_PIN_ST(u) asm volatile("st.shared.f32 [%0+%2], %1;" :: "r"(dsta), "f"(echo[u]), "n"((u) * 144) : "memory");
                        MCRT_PIN_ST(0) MCRT_PIN_ST(1) MCRT_PIN_ST(2) MCRT_PIN_ST(3) MCRT_PIN_ST(4) MCRT_PIN_ST(5) MCRT_PIN_ST(6)
#undef MCRT_PIN_ST
                    } else
#endif
                    {
                    float* dst = my_col + slot0 * stride;
                    if (cont) dst[-stride] = cur_acc;
#pragma unroll
                    for (int u = 0; u < MCRT_WIN_UNROLL - 1; u++) dst[u * stride] = echo[u];
                    }
                    cur_acc = echo[MCRT_WIN_UNROLL - 1];
                } else {
                    add_row(echo[0], row0);
                    for (int r = written > base ? written : base; r < row0; r++) my_col[MCRT_WIN_SLOT(r) * stride] = 0.0f;   // gap (time jumped ahead)
                    // rows row0+1 .. row0+U-1 receive exactly one echo each: close every row straight into the column
                    if (slot0 + MCRT_WIN_UNROLL <= MCRT_WIN_RING) {                                 // no wrap inside the block
                        float* dst = my_col + slot0 * stride;
#pragma unroll
                        for (int u = 1; u < MCRT_WIN_UNROLL; u++) {
                            *dst = cur_acc;
                            dst += stride;
                            cur_acc = echo[u];
                        }
                    } else {
#pragma unroll
                        for (int u = 1; u < MCRT_WIN_UNROLL; u++) {
                            my_col[MCRT_WIN_SLOT(row0 + u - 1) * stride] = cur_acc;
                            cur_acc = echo[u];
                        }
                    }
                }
                written = row0 + MCRT_WIN_UNROLL - 1;
                cur_row = row0 + MCRT_WIN_UNROLL - 1;
                n_safe -= MCRT_WIN_UNROLL; remaining -= MCRT_WIN_UNROLL;
                my_steps += MCRT_WIN_UNROLL;
            }
#if MCRT_WIN_TAIL_BLOCK
            // ---- tail block: the last m < MCRT_WIN_UNROLL steps of a segment that ends by its step budget (remaining == n_safe), taken like an
            // unrolled block -- all volume loads in flight together, rows row0 + u under the same guard -- instead of m checked steps with an
            // exposed load each.  The warp pays for the LONGEST tail among its lanes, so up to seven serial checked steps become one block.  The
            // steps beyond m are computed and dropped: indices wrap inside the volume, and the march state (point, time, intensity) is dead once
            // the segment has ended -- the closing echo uses end_micros and the next segment reloads everything.
            if (n_safe >= MCRT_WIN_TAIL_BLOCK && n_safe < MCRT_WIN_UNROLL && remaining == n_safe && FMADIV && fma_ok) {
                const double rowd0 = time_elapsed * inv_row_period;
                const int row0 = __double2int_rd(rowd0);
                const double f0 = rowd0 - (double)row0;
                if (fast_rows && f0 >= 1e-6 && f0 <= block_safe_hi && row0 < wend && row0 + MCRT_WIN_UNROLL <= rows && row0 >= base && row0 >= cur_row) {
                    const int m = n_safe;
                    uint32_t idx[MCRT_WIN_UNROLL];
#pragma unroll
                    for (int u = 0; u < MCRT_WIN_UNROLL; u++) { idx[u] = voxel_linear_fma(point, vres, inv_vres); point = v_add(point, delta_step); }
                    float2 vox[MCRT_WIN_UNROLL];
#pragma unroll
                    for (int u = 0; u < MCRT_WIN_UNROLL; u++) vox[u] = __ldg(&volume[idx[u]]);
                    float echo[MCRT_WIN_UNROLL];
#pragma unroll
                    for (int u = 0; u < MCRT_WIN_UNROLL; u++) {
                        const float scattering = vox[u].y >= m_mu1 ? vox[u].x * m_sigma + m_mu0 : 0.0f;
                        echo[u] = intensity * scattering;
                        intensity *= decay;
                    }
                    add_row(echo[0], row0);
                    for (int r = written > base ? written : base; r < row0; r++) my_col[MCRT_WIN_SLOT(r) * stride] = 0.0f;   // gap (time jumped ahead)
#pragma unroll
                    for (int u = 1; u < MCRT_WIN_UNROLL; u++) {
                        if (u < m) {
                            my_col[MCRT_WIN_SLOT(row0 + u - 1) * stride] = cur_acc;
                            cur_acc = echo[u];
                        }
                    }
                    written = row0 + m - 1;
                    cur_row = row0 + m - 1;
                    my_steps += m;
                    remaining = 0; n_safe = 0;
                }
            }
#endif
            // ---- single checked steps: window edges, guard bands, the tail of the segment ----
            bool seg_done = false;
            while (true) {
                if (!(remaining > 0 && time_elapsed < max_travel_time)) { seg_done = true; break; }
                if (n_safe >= MCRT_WIN_UNROLL && (!FMADIV || fma_ok)) {
                    // back to the block path as soon as a whole block fits this window again
                    const int row_now = __double2int_rd(time_elapsed * inv_row_period);
                    if (fast_rows && row_now < wend && row_now + MCRT_WIN_UNROLL <= rows && row_now >= base) {
                        const double f0 = time_elapsed * inv_row_period - (double)row_now;
                        if (fast_rows && f0 >= 1e-6 && f0 <= block_safe_hi && row_now >= cur_row) break;
                    }
#if MCRT_WIN_EDGE_SKIP
                    else if (fast_rows && row_now >= wend && row_now < rows) {
                        // the common way out of the block loop: the next echo lies in a later window.  Under the guard on f0 row_now IS the row
                        // try_echo would compute, so the checked step below would load its voxel, form the echo and learn that nothing can be
                        // consumed; say so here (kept out of the block loop: any state added there costs registers the kernel does not have)
                        const double f0 = time_elapsed * inv_row_period - (double)row_now;
                        if (f0 >= 1e-6 && f0 <= 1.0 - 1e-6) { if (TREE) pend_row = row_now; waiting = true; break; }
                    }
#endif
                }
                const float2 vox = __ldg(&volume[FMADIV ? (fma_ok ? voxel_linear_fma(point, vres, inv_vres) : voxel_linear_exact(point.x, point.y, point.z, vres))
                                                         : voxel_linear(point, vres, inv_vres)]);
                const float scattering = vox.y >= m_mu1 ? vox.x * m_sigma + m_mu0 : 0.0f;
                if (!try_echo(intensity * scattering, time_elapsed)) { waiting = true; break; }
                point = v_add(point, delta_step);
                time_elapsed = time_elapsed + time_step;
                intensity *= decay;
                remaining--;
                if (n_safe > 0) n_safe--;
                my_steps++;
            }
            if (seg_done) {
                // main.cpp:139: the closing echo of the segment
                if (!try_echo(end_echo, end_micros)) { waiting = true; remaining = 0; n_safe = 0; }
                else in_seg = false;
            }
        }
        // close my column of this window: pending row (if it lies in the window) and zeros up to the window end
        if (t < T) {
            if (cur_row >= base && cur_row < wend) {
                for (int r = written; r < cur_row; r++) my_col[MCRT_WIN_SLOT(r) * stride] = 0.0f;
                my_col[MCRT_WIN_SLOT(cur_row) * stride] = cur_acc;
                written = cur_row + 1;
                cur_row = -1; cur_acc = 0.0f;
            }
            if (written < base) written = base;
            for (int r = written; r < wend; r++) my_col[MCRT_WIN_SLOT(r) * stride] = 0.0f;
            if (written < wend) written = wend;
        }
        group_sync();
        // sample reduction in sample order (= k_reduce_samples): thread r of the group owns row base + r of the window and walks
        // the group's scanlines, so consecutive lanes store consecutive rows (coalesced) and no index arithmetic is needed
        const int wrows = wend - base;
        if (t < wrows) {
            const float* rowp = s_win + MCRT_WIN_SLOT(base + t) * stride;
            float* dst = rf + (size_t)scanline0 * rf_pitch + base + t;
            const int g_end = n_scanlines - scanline0 < G ? n_scanlines - scanline0 : G;
            if (SCT > 0 && (SCT & 3) == 0) {
#pragma unroll
                for (int g = 0; g < 32 / (SCT > 0 ? SCT : 32); g++) {
                    if (g < g_end) {
                        const float4* s4 = reinterpret_cast<const float4*>(rowp + g * SCT);
                        float4 v = s4[0];
                        float sum = v.x; sum += v.y; sum += v.z; sum += v.w;
#pragma unroll
                        for (int q = 1; q < (SCT > 0 ? SCT : 4) / 4; q++) { v = s4[q]; sum += v.x; sum += v.y; sum += v.z; sum += v.w; }
                        if (TREE) dst[(size_t)g * rf_pitch] += sum;           // rounds accumulate (same warp, program order)
                        else dst[(size_t)g * rf_pitch] = sum;
                    }
                }
            } else {
                for (int g = 0; g < g_end; g++) {
                    const float* src = rowp + g * S;
                    float sum;
                    if ((S & 3) == 0) {
                        const float4* s4 = reinterpret_cast<const float4*>(src);
                        float4 v = s4[0];
                        sum = v.x; sum += v.y; sum += v.z; sum += v.w;
                        for (int q = 1; q < (S >> 2); q++) { v = s4[q]; sum += v.x; sum += v.y; sum += v.z; sum += v.w; }
                    } else {
                        sum = src[0];
                        for (int s = 1; s < S; s++) sum += src[s];
                    }
                    dst[(size_t)g * rf_pitch] = sum;
                }
            }
        }
        group_sync();
    }
    }   // rounds (TREE), a single pass otherwise
    if (steps_total) {
        for (int off = 16; off > 0; off >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, off);
        if ((threadIdx.x & 31) == 0 && my_steps) atomicAdd(steps_total, my_steps);
    }
}

__global__ void __launch_bounds__(256) k_reduce_samples(const float* __restrict__ columns, const int64_t n_pixels, const int samples,
                                                       float* __restrict__ rf, const int rows, const int rf_pitch)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += (int64_t)gridDim.x * blockDim.x) {
        const float* src = columns + i * samples;
        float sum;
        if ((samples & 3) == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4 v = __ldg(&s4[0]);
            sum = v.x; sum += v.y; sum += v.z; sum += v.w;
            for (int t = 1; t < (samples >> 2); t++) {
                v = __ldg(&s4[t]);
                sum += v.x; sum += v.y; sum += v.z; sum += v.w;
            }
        } else {
            sum = __ldg(&src[0]);
            for (int t = 1; t < samples; t++) sum += __ldg(&src[t]);
        }
        if (rf_pitch == rows) rf[i] = sum;
        else { const int64_t sl = i / rows; rf[sl * rf_pitch + (i - sl * rows)] = sum; }
    }
}

// ------------------------------------------------------------------------------------------------
// rf_image::convolve (rfimage.h:93-123).  Pass 1 (axial, along a scanline) and pass 2 (lateral,
// across scanlines) keep the reference's forward-looking taps, its sequential fp32 sum order and
// its untouched borders (B-9): rows [0,Ka) and [rows-Ka,rows), columns [0,Kl/2) and [cols-Kl,cols)
// keep the raw samples.
// ------------------------------------------------------------------------------------------------
#define MCRT_MAX_TAPS 256

// ---- long-scanline path (rows that do not fit the fused kernel's shared memory, BASELINE config 5) ----
// Register-blocked so each loaded sample feeds MCRT_PSF_R outputs: 2 + ~1/R instructions per tap and
// output instead of a load + multiply + add + loop overhead per tap.
#define MCRT_PSF_R 8                  // consecutive scanlines per thread (lateral pass)
#define MCRT_PSF_RA 8                 // consecutive rows per thread (axial pass).  16 cuts the instruction count by 12 % (2 LDS + 1 branch per
                                      // 32 tap applications) but leaves a third of every CTA's threads idle on 8333-row scanlines: 0.51 vs 0.44 ms
#define MCRT_PSF_CHUNK (256 * MCRT_PSF_RA)   // most output rows per CTA of k_psf_axial

// logical row index within the staged chunk -> shared-memory word: one pad word per MCRT_PSF_RA words, so
// the lanes of a warp (whose windows start MCRT_PSF_RA apart) hit distinct banks
__host__ __device__ __forceinline__ int psf_pad(int i) { return i + i / MCRT_PSF_RA; }

// Axial pass (rfimage.h:97-108).  CTA = one piece of one scanline, staged through shared memory (coalesced 128 B
// loads); thread t owns MCRT_PSF_RA consecutive outputs and slides a register window of that many samples along the taps.  For every
// output the taps are applied in order k = 0.. with separate multiply and add (the reference's sequential fp32 sum).
// Rows outside [ka, rows-ka) are never read downstream.
__global__ void __launch_bounds__(256) k_psf_axial(const float* __restrict__ in, const int rows, const float* __restrict__ taps, const int ka,
                                                  float* __restrict__ out, const int in_pitch, const int chunk_rows)
{
    // chunk_rows <= MCRT_PSF_CHUNK, a multiple of MCRT_PSF_RA: the scanline is cut into EVEN pieces (round 1: 4 x 2048 + 141 rows
    // for 8333, whose fifth CTA staged and synchronised for 7 % of a chunk)
    extern __shared__ float s_row[];                       // psf_pad(chunk_rows + ka + MCRT_PSF_RA) words
    __shared__ float s_taps[MCRT_MAX_TAPS + MCRT_PSF_RA];
    const int tid = threadIdx.x;
    const int chunk0 = blockIdx.x * chunk_rows;            // first output row of this CTA
    const float* src = in + (size_t)blockIdx.y * in_pitch;
    float* dst = out + (size_t)blockIdx.y * rows;
    for (int i = tid; i < ka; i += blockDim.x) s_taps[i] = taps[i];
    const int n_stage = chunk_rows + ka + MCRT_PSF_RA;
    for (int i = tid; i < n_stage; i += blockDim.x) {
        const int r = chunk0 + i;
        s_row[psf_pad(i)] = r < rows ? __ldg(&src[r]) : 0.0f;
    }
    __syncthreads();
    const int l0 = tid * MCRT_PSF_RA;                      // first output row of this thread, chunk-local
    const int r0 = chunk0 + l0;
    if (l0 >= chunk_rows || r0 >= rows - ka || r0 + MCRT_PSF_RA <= ka) return;
    float acc[MCRT_PSF_RA], win[MCRT_PSF_RA];
#pragma unroll
    for (int j = 0; j < MCRT_PSF_RA; j++) { acc[j] = 0.0f; win[j] = s_row[psf_pad(l0 + j)]; }
    // the window refill of tap k reads logical row l0 + k + R = R (tid + 1) + k, i.e. padded word (R + 1) (tid + 1) + (R + 1) (k / R) + k % R:
    // a pointer that advances R + 1 words per group of R taps plus a compile-time offset -- no address arithmetic in the tap loop
    const float* wp = s_row + psf_pad(l0 + MCRT_PSF_RA);
    const float* tp = s_taps;
    for (int k0 = 0; k0 < ka; k0 += MCRT_PSF_RA, wp += MCRT_PSF_RA + 1, tp += MCRT_PSF_RA) {
#pragma unroll
        for (int kk = 0; kk < MCRT_PSF_RA; kk++) {
            if (k0 + kk < ka) {
                const float t = tp[kk];
                // window invariant: win[(kk + j) % R] == sample at row l0 + k + j
#pragma unroll
                for (int j = 0; j < MCRT_PSF_RA; j++) acc[j] += win[(kk + j) & (MCRT_PSF_RA - 1)] * t;
                win[kk] = wp[kk];                                  // slot kk held row l0+k, now row l0+k+R
            }
        }
    }
#pragma unroll
    for (int j = 0; j < MCRT_PSF_RA; j++) {
        const int r = r0 + j;
        if (r >= ka && r < rows - ka) dst[r] = acc[j];
    }
}

// The same pass with the tap loop resolved at compile time.  NG = number of groups of MCRT_PSF_RA taps (ka <= NG * MCRT_PSF_RA; only the
// LAST group checks k < ka, a uniform branch per tail tap), taps = kernel parameters, so every multiply takes its tap from the constant
// bank and every window refill is an LDS at an immediate offset: per 8 outputs and tap 16 FP instructions + 1 LDS, nothing else
// (the run-time loop above spends another LDS for the tap, a predicate and the loop arithmetic: 1 550 instead of ~1 100 warp
// instructions per 8 x 63 tap applications).  Same arithmetic and order: bit-identical.
struct LongTaps { float t[64]; };
template <int NG>
__global__ void __launch_bounds__(256) k_psf_axial_ct(const float* __restrict__ in, const int rows, const __grid_constant__ LongTaps taps, const int ka,
                                                     float* __restrict__ out, const int in_pitch, const int chunk_rows)
{
    extern __shared__ float s_row[];                       // psf_pad(chunk_rows + ka + MCRT_PSF_RA) words
    const int tid = threadIdx.x;
    const int chunk0 = blockIdx.x * chunk_rows;
    const float* src = in + (size_t)blockIdx.y * in_pitch;
    float* dst = out + (size_t)blockIdx.y * rows;
    const int n_stage = chunk_rows + NG * MCRT_PSF_RA + MCRT_PSF_RA;
    for (int i = tid; i < n_stage; i += 256) {
        const int r = chunk0 + i;
        s_row[psf_pad(i)] = r < rows ? __ldg(&src[r]) : 0.0f;
    }
    __syncthreads();
    const int l0 = tid * MCRT_PSF_RA;
    const int r0 = chunk0 + l0;
    if (l0 >= chunk_rows || r0 >= rows - ka || r0 + MCRT_PSF_RA <= ka) return;
    float acc[MCRT_PSF_RA], win[MCRT_PSF_RA];
    const float* wp = s_row + psf_pad(l0);
#pragma unroll
    for (int j = 0; j < MCRT_PSF_RA; j++) { acc[j] = 0.0f; win[j] = wp[j]; }
#pragma unroll
    for (int g = 0; g < NG; g++) {
#pragma unroll
        for (int kk = 0; kk < MCRT_PSF_RA; kk++) {
            const int k = g * MCRT_PSF_RA + kk;
            if (g < NG - 1 || k < ka) {
                const float t = taps.t[k];
#pragma unroll
                for (int j = 0; j < MCRT_PSF_RA; j++) acc[j] += win[(kk + j) & (MCRT_PSF_RA - 1)] * t;
                win[kk] = wp[(g + 1) * (MCRT_PSF_RA + 1) + kk];       // logical row l0 + k + R
            }
        }
    }
#pragma unroll
    for (int j = 0; j < MCRT_PSF_RA; j++) {
        const int r = r0 + j;
        if (r >= ka && r < rows - ka) dst[r] = acc[j];
    }
}

// Lateral pass (rfimage.h:111-122) + untouched borders (B-9).  Thread = one row, MCRT_PSF_R consecutive
// scanlines; lanes walk consecutive rows, so every load/store is coalesced and no shared memory is needed.
// BYROW: depth-dependent lateral PSF, tap k of RF row r = taps_by_row[k * rows + r].
template <bool BYROW>
__global__ void __launch_bounds__(256) k_psf_lateral(const float* __restrict__ raw, const float* __restrict__ axial_buf, const int cols,
                                                    const int rows, const float* __restrict__ taps, const int ka, const int kl,
                                                    const int col_offset, const int cols_total, float* __restrict__ out,
                                                    const float* __restrict__ taps_by_row, const int raw_pitch)
{
    __shared__ float s_taps[MCRT_MAX_TAPS];
    for (int i = threadIdx.x; i < kl; i += blockDim.x) s_taps[i] = taps[i];
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int c0 = blockIdx.y * MCRT_PSF_R;
    const size_t img = (size_t)blockIdx.z * cols * rows;
    const bool row_ok = r >= ka && r < rows - ka;
    // border rule in GLOBAL scanline indices (a scanline-block run holds scanlines col_offset.. of cols_total);
    // does any of the 8 scanlines get convolved at all?
    const bool any = row_ok && (col_offset + c0 + MCRT_PSF_R > kl / 2) && (col_offset + c0 < cols_total - kl);
    float acc[MCRT_PSF_R], win[MCRT_PSF_R];
    if (any) {
        // scanline c0 + j of this row; the pointer walks one scanline (rows floats) per tap
        const float* p = axial_buf + img + (size_t)c0 * rows + r;
        const int c_left = cols - c0;                          // scanlines that exist from c0 on
#pragma unroll
        for (int j = 0; j < MCRT_PSF_R; j++) { acc[j] = 0.0f; win[j] = j < c_left ? __ldg(p + (size_t)j * rows) : 0.0f; }
        p += (size_t)MCRT_PSF_R * rows;
        const float* tr = BYROW ? taps_by_row + r : nullptr;
        for (int k0 = 0; k0 < kl; k0 += MCRT_PSF_R) {
#pragma unroll
            for (int kk = 0; kk < MCRT_PSF_R; kk++) {
                const int k = k0 + kk;
                if (k < kl) {
                    const float t = BYROW ? __ldg(tr + (size_t)k * rows) : s_taps[k];
                    // window invariant: win[(kk + j) & 7] == axial value of scanline c0 + k + j
#pragma unroll
                    for (int j = 0; j < MCRT_PSF_R; j++) acc[j] += win[(kk + j) & (MCRT_PSF_R - 1)] * t;
                    win[kk] = k + MCRT_PSF_R < c_left ? __ldg(p) : 0.0f;
                    p += rows;
                }
            }
        }
    }
    float* o = out + img + (size_t)c0 * rows + r;
    const float* rw = raw + ((size_t)blockIdx.z * cols + c0) * raw_pitch + r;
#pragma unroll
    for (int j = 0; j < MCRT_PSF_R; j++) {
        const int c = c0 + j;
        if (c >= cols) break;
        o[(size_t)j * rows] = (any && col_offset + c >= kl / 2 && col_offset + c < cols_total - kl) ? acc[j] : __ldg(rw + (size_t)j * raw_pitch);
    }
}

// The lateral pass with the tap loop resolved at compile time (NG groups of MCRT_PSF_RL taps, only the last group checks k < kl; taps in
// the constant bank) and 16 scanlines per thread: every loaded value feeds 16 outputs (32 FP instructions per LDG + pointer step, and
// (16 + kl - 1) / 16 = 2.9 reads of every input value through L2 instead of 4.75 with 8 scanlines).  Interior scanline groups -- all
// 16 outputs convolved, all kl + 15 inputs inside the image -- carry no guards at all.  Same arithmetic and order: bit-identical.
#define MCRT_PSF_RL 16
template <int NG, bool GUARD>
__device__ __forceinline__ void lateral_ct_body(const float* __restrict__ p, const int rows, const LongTaps& taps, const int kl, const int c_left,
                                                float (&acc)[MCRT_PSF_RL])
{
    float win[MCRT_PSF_RL];
#pragma unroll
    for (int j = 0; j < MCRT_PSF_RL; j++) { acc[j] = 0.0f; win[j] = (!GUARD || j < c_left) ? __ldg(p + (size_t)j * rows) : 0.0f; }
    p += (size_t)MCRT_PSF_RL * rows;
#pragma unroll
    for (int g = 0; g < NG; g++) {
#pragma unroll
        for (int kk = 0; kk < MCRT_PSF_RL; kk++) {
            const int k = g * MCRT_PSF_RL + kk;
            if (g < NG - 1 || k < kl) {
                const float t = taps.t[k];
#pragma unroll
                for (int j = 0; j < MCRT_PSF_RL; j++) acc[j] += win[(kk + j) & (MCRT_PSF_RL - 1)] * t;
                // the refill of tap k is first used by tap k + 1: not needed after the last tap
                if (g < NG - 1 || k + 1 < kl) win[kk] = (!GUARD || k + MCRT_PSF_RL < c_left) ? __ldg(p) : 0.0f;
                p += rows;
            }
        }
    }
}

template <int NG>
__global__ void __launch_bounds__(256) k_psf_lateral_ct(const float* __restrict__ raw, const float* __restrict__ axial_buf, const int cols,
                                                       const int rows, const __grid_constant__ LongTaps taps, const int ka, const int kl,
                                                       const int col_offset, const int cols_total, float* __restrict__ out, const int raw_pitch)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int c0 = blockIdx.y * MCRT_PSF_RL;
    const size_t img = (size_t)blockIdx.z * cols * rows;
    const bool row_ok = r >= ka && r < rows - ka;
    const int c_left = cols - c0;
    float* o = out + img + (size_t)c0 * rows + r;
    const float* p = axial_buf + img + (size_t)c0 * rows + r;
    float acc[MCRT_PSF_RL];
    // interior group (uniform per CTA): every scanline of the group is convolved and every input scanline exists
    if (col_offset + c0 >= kl / 2 && col_offset + c0 + MCRT_PSF_RL <= cols_total - kl && c_left >= MCRT_PSF_RL + kl - 1) {
        if (row_ok) {
            lateral_ct_body<NG, false>(p, rows, taps, kl, c_left, acc);
#pragma unroll
            for (int j = 0; j < MCRT_PSF_RL; j++) o[(size_t)j * rows] = acc[j];
            return;
        }
        const float* rw = raw + ((size_t)blockIdx.z * cols + c0) * raw_pitch + r;
#pragma unroll
        for (int j = 0; j < MCRT_PSF_RL; j++) o[(size_t)j * rows] = __ldg(rw + (size_t)j * raw_pitch);
        return;
    }
    // border groups: the rule in GLOBAL scanline indices (a scanline-block run holds scanlines col_offset.. of cols_total)
    const bool any = row_ok && (col_offset + c0 + MCRT_PSF_RL > kl / 2) && (col_offset + c0 < cols_total - kl);
    if (any) lateral_ct_body<NG, true>(p, rows, taps, kl, c_left, acc);
    const float* rw = raw + ((size_t)blockIdx.z * cols + c0) * raw_pitch + r;
#pragma unroll
    for (int j = 0; j < MCRT_PSF_RL; j++) {
        const int c = c0 + j;
        if (c >= cols) break;
        o[(size_t)j * rows] = (any && col_offset + c >= kl / 2 && col_offset + c < cols_total - kl) ? acc[j] : __ldg(rw + (size_t)j * raw_pitch);
    }
}

// ------------------------------------------------------------------------------------------------
// rf_image::envelope (rfimage.h:54-91).  A sample i in [1, rows-2] is a peak iff I[i-1] < I[i] and not
// I[i] < I[i+1] (the sequential `ascending` flag reduces to this because peak detection only ever reads
// samples that have not been rewritten yet).  Between consecutive peaks p < q the output is
// lerp(last, |I[q]|) with last = I[0] (signed) for the virtual first peak and |I[p]| otherwise; samples
// from the last peak on keep their raw value.
// Long-scanline versions below (round 1's k_peak_masks + k_envelope_lerp pair is gone).
// ------------------------------------------------------------------------------------------------
// Long scanlines, round 2: ONE kernel (round 1: k_peak_masks + a fully parallel k_envelope_lerp that paid a 64-bit division and two
// mask scans per SAMPLE, 190 instruction slots per pixel).  Same arithmetic as k_post_fused's envelope: bit-identical.
#define MCRT_ENV_WARPS 8
__global__ void __launch_bounds__(MCRT_ENV_WARPS * 32) k_envelope_stream(const float* __restrict__ in, const int64_t n_scanlines, const int rows,
                                                                        const int words, float* __restrict__ out)
{
    // one CTA per scanline: the 32-row chunks are independent once every chunk knows the last peak before it and the first peak
    // after it, so the warps take chunks round-robin (many independent loads in flight) and only the tiny scan over the chunk
    // masks in between is serial
    extern __shared__ unsigned s_env[];                    // masks[words], next[words], last[words]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned* mask = s_env;
    int* nxt = reinterpret_cast<int*>(mask + words);
    int* lst = nxt + words;
    for (int64_t sl = blockIdx.x; sl < n_scanlines; sl += gridDim.x) {
        const float* I = in + sl * rows;
        float* O = out + sl * rows;
        for (int ch = w; ch < words; ch += MCRT_ENV_WARPS) {
            const int i = (ch << 5) + lane;
            const float v = i < rows ? __ldg(&I[i]) : 0.0f;
            const float vm = (i >= 1 && i < rows) ? __ldg(&I[i - 1]) : 0.0f;
            const float vp = (i + 1 < rows) ? __ldg(&I[i + 1]) : 0.0f;
            const bool peak = (i >= 1) && (i + 1 < rows) && (vm < v) && !(v < vp);
            const unsigned m = __ballot_sync(0xffffffffu, peak);
            if (lane == 0) mask[ch] = m;
        }
        __syncthreads();
        // first peak after chunk ch / last peak before chunk ch (0 = the virtual first peak): peaks are a few samples apart in RF
        // data, so the scans almost always stop at the neighbouring word
        for (int ch = threadIdx.x; ch < words; ch += blockDim.x) {
            int q = rows;
            for (int c = ch + 1; c < words; c++) { const unsigned m = mask[c]; if (m) { q = (c << 5) + (__ffs(m) - 1); break; } }
            int p = 0;
            for (int c = ch - 1; c >= 0; c--) { const unsigned m = mask[c]; if (m) { p = (c << 5) + (31 - __clz(m)); break; } }
            nxt[ch] = q; lst[ch] = p;
        }
        __syncthreads();
        const float first = __ldg(&I[0]);
        for (int ch = w; ch < words; ch += MCRT_ENV_WARPS) {
            const unsigned m = mask[ch];
            const int i = (ch << 5) + lane;
            const unsigned le = m & (0xffffffffu >> (31 - lane));
            const int p = le ? (ch << 5) + (31 - __clz(le)) : lst[ch];
            const unsigned gt = lane == 31 ? 0u : (m & (0xffffffffu << (lane + 1)));
            const int q = gt ? (ch << 5) + (__ffs(gt) - 1) : nxt[ch];
            if (i < rows) {
                float r = __ldg(&I[i]);
                if (q < rows) {
                    const float last = (p == 0) ? first : fabsf(__ldg(&I[p]));
                    const float new_peak = fabsf(__ldg(&I[q]));
                    const float alpha = ((float)i - (float)p) / ((float)q - (float)p);
                    r = last * (1 - alpha) + new_peak * alpha;
                }
                O[i] = r;
            }
        }
        __syncthreads();
    }
}

// Round 2, second step (k_envelope_tiles): a scanline is cut into tiles of 32 chunks (1024 rows), ONE WARP PER TILE, a CTA per scanline.
// The warp loads its tile with 32 independent coalesced loads and keeps it in registers; the peak test takes its neighbours from
// the adjacent lanes (shuffles), a peak value inside the sample's own chunk is a shuffle as well, and the nearest peaks outside the
// chunk travel as warp-uniform (position, value) pairs: a backward sweep over the tile's chunk masks gives every chunk the first
// peak after it, a running pair the last peak before it, and the tiles exchange their first / last peak through a few words of
// shared memory -- ONE barrier per scanline, no per-sample mask scans, no re-read of the scanline.
// alpha = (i - p) / (q - p) is the 3-instruction Markstein division from a reciprocal table (div_small_int, exhaustively equal to the
// IEEE quotient for gaps up to 2048 rows; larger gaps take the IEEE division).  Same arithmetic as k_envelope_stream: bit-identical.
#define MCRT_ENV_RCP 2048
#ifndef MCRT_ENVT_CH
#define MCRT_ENVT_CH 32                                   // chunks per warp tile
#endif
#ifndef MCRT_ENVT_MINB
#define MCRT_ENVT_MINB 3                                  // resident CTAs per SM the register allocation aims at (<= 640-thread variants)
#endif
__device__ __forceinline__ float div_small_int(float a, float b, float y);
template <int MAXT>
__global__ void __launch_bounds__(MAXT, (MAXT <= 640 ? MCRT_ENVT_MINB : 1)) k_envelope_tiles(const float* __restrict__ in, const int64_t n_scanlines, const int rows,
                                                        float* __restrict__ out)
{
    extern __shared__ unsigned s_env[];    // rcp[n_rcp + 1] | per warp: 7 x 32 words (below) | 2 x per warp: first pos / val, last pos / val
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n_rcp = rows < MCRT_ENV_RCP ? rows : MCRT_ENV_RCP;
    float* rcp = reinterpret_cast<float*>(s_env);
    unsigned* wmask = s_env + (n_rcp + 1) + w * (7 * MCRT_ENVT_CH);      // peak mask of chunk c
    float* wfv = reinterpret_cast<float*>(wmask + MCRT_ENVT_CH);        // |value| of the chunk's first peak
    float* wlv = wfv + MCRT_ENVT_CH;                                    // |value| of the chunk's last peak
    int* wnq = reinterpret_cast<int*>(wlv + MCRT_ENVT_CH);              // first peak after chunk c: row (rows: none) ...
    float* wnv = reinterpret_cast<float*>(wnq + MCRT_ENVT_CH);          // ... and |value|
    int* wlp = reinterpret_cast<int*>(wnv + MCRT_ENVT_CH);              // last peak before chunk c: row (0: the virtual first peak) ...
    float* wlq = reinterpret_cast<float*>(wlp + MCRT_ENVT_CH);          // ... and value
    int* summ = reinterpret_cast<int*>(s_env + (n_rcp + 1) + nw * (7 * MCRT_ENVT_CH));   // [2][nw][4]
    for (int i = threadIdx.x; i <= n_rcp; i += blockDim.x) rcp[i] = i > 0 ? 1.0f / (float)i : 0.0f;
    const unsigned le_mask = 0xffffffffu >> (31 - lane);                  // bits <= lane
    const unsigned gt_mask = lane == 31 ? 0u : (0xffffffffu << (lane + 1));
    const unsigned lt_mask = (1u << lane) - 1u;
    const int tile0 = w * (32 * MCRT_ENVT_CH);
    int parity = 0;
    for (int64_t sl = blockIdx.x; sl < n_scanlines; sl += gridDim.x, parity ^= 1) {
        const float* I = in + sl * rows + tile0;
        float* O = out + sl * rows + tile0;
        // ---- the tile -> registers; per chunk: peak mask and the values of its first / last peak -> shared memory ----
        float v[MCRT_ENVT_CH];
#pragma unroll
        for (int c = 0; c < MCRT_ENVT_CH; c++) v[c] = tile0 + 32 * c + lane < rows ? __ldg(I + 32 * c + lane) : 0.0f;
        const float left = (lane == 0 && tile0 >= 1) ? __ldg(I - 1) : 0.0f;
        const float right = (lane == 31 && tile0 + 32 * MCRT_ENVT_CH < rows) ? __ldg(I + 32 * MCRT_ENVT_CH) : 0.0f;
#pragma unroll
        for (int c = 0; c < MCRT_ENVT_CH; c++) {
            const int i = tile0 + 32 * c + lane;
            float vm = __shfl_up_sync(0xffffffffu, v[c], 1), vp = __shfl_down_sync(0xffffffffu, v[c], 1);
            const float prev_last = c > 0 ? __shfl_sync(0xffffffffu, v[c > 0 ? c - 1 : 0], 31) : left;
            const float next_first = c + 1 < MCRT_ENVT_CH ? __shfl_sync(0xffffffffu, v[c + 1 < MCRT_ENVT_CH ? c + 1 : c], 0) : right;
            if (lane == 0) vm = prev_last;
            if (lane == 31) vp = next_first;
            const bool peak = (i >= 1) && (i + 1 < rows) && (vm < v[c]) && !(v[c] < vp);
            const unsigned m = __ballot_sync(0xffffffffu, peak);
            // (an empty mask shuffles from lanes 31 / 0: the values are never used)
            const float fv = __shfl_sync(0xffffffffu, v[c], (__ffs(m) - 1) & 31), lv = __shfl_sync(0xffffffffu, v[c], (31 - __clz(m)) & 31);
            if (lane == 0) { wmask[c] = m; wfv[c] = fabsf(fv); wlv[c] = fabsf(lv); }
        }
        __syncwarp();
        // ---- lane c = chunk c: the nearest non-empty chunks of the tile on either side ----
        const unsigned mc = wmask[lane];
        const unsigned B = __ballot_sync(0xffffffffu, mc != 0u);
        const unsigned after = B & gt_mask, before = B & lt_mask;
        const int na = __ffs(after) - 1, nb = 31 - __clz(before);
        const int nq_in = after ? tile0 + 32 * na + (__ffs(wmask[na & 31]) - 1) : -1;
        const float nv_in = wfv[na & 31];
        const int lp_in = before ? tile0 + 32 * nb + (31 - __clz(wmask[nb & 31])) : -1;
        const float lv_in = wlv[nb & 31];
        // the tile's first / last peak for the other tiles
        int* my = summ + (parity * nw + w) * 4;
        if (lane == 0) {
            const int f = __ffs(B) - 1, l = 31 - __clz(B);
            my[0] = B ? tile0 + 32 * f + (__ffs(wmask[f & 31]) - 1) : -1; my[1] = __float_as_int(wfv[f & 31]);
            my[2] = B ? tile0 + 32 * l + (31 - __clz(wmask[l & 31])) : -1; my[3] = __float_as_int(wlv[l & 31]);
        }
        __syncthreads();
        // ---- nearest peaks outside the tile ----
        int rq = rows; float rv = 0.0f;                                      // first peak after the tile (rows: none)
        for (int t = w + 1; t < nw; t++) { const int* o = summ + (parity * nw + t) * 4; if (o[0] >= 0) { rq = o[0]; rv = __int_as_float(o[1]); break; } }
        int lp = 0; float lval = __ldg(in + sl * rows);                        // last peak before the tile (0: the virtual first peak (0, I[0]))
        for (int t = w - 1; t >= 0; t--) { const int* o = summ + (parity * nw + t) * 4; if (o[2] >= 0) { lp = o[2]; lval = __int_as_float(o[3]); break; } }
        wnq[lane] = nq_in >= 0 ? nq_in : rq; wnv[lane] = nq_in >= 0 ? nv_in : rv;
        wlp[lane] = lp_in >= 0 ? lp_in : lp; wlq[lane] = lp_in >= 0 ? lv_in : lval;
        __syncwarp();
        // ---- interpolate ----
#pragma unroll
        for (int c = 0; c < MCRT_ENVT_CH; c++) {
            const int base = tile0 + 32 * c;
            if (base < rows) {                                               // warp-uniform
                const int i = base + lane;
                const unsigned m = wmask[c];
                const unsigned le = m & le_mask, gt = m & gt_mask;
                const int jl = 31 - __clz(le), jg = __ffs(gt) - 1;           // (garbage lanes when le / gt are empty: masked below)
                const float vl = __shfl_sync(0xffffffffu, v[c], jl & 31), vg = __shfl_sync(0xffffffffu, v[c], jg & 31);
                const int p = le ? base + jl : wlp[c];
                const float last = le ? fabsf(vl) : wlq[c];
                const int q = gt ? base + jg : wnq[c];
                const float new_peak = gt ? fabsf(vg) : wnv[c];
                float r = v[c];
                if (q < rows) {
                    const int d = q - p;
                    const float fa = (float)(i - p), fd = (float)d;          // == (float)i - (float)p, (float)q - (float)p: exact integers
                    const float alpha = d <= MCRT_ENV_RCP ? div_small_int(fa, fd, rcp[d <= MCRT_ENV_RCP ? d : 0]) : fa / fd;
                    r = last * (1 - alpha) + new_peak * alpha;
                }
                if (i < rows) O[32 * c + lane] = r;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Fused post-processing for scanlines that fit shared memory (the reference's 465-row images):
// one CTA stages TC output scanlines plus the Kl-1 halo scanlines to their right (whole columns, so
// the forward-looking axial taps need no row halo), runs the axial pass and the lateral pass out of
// shared memory and finishes with the envelope, one warp per scanline -- the image is read from HBM
// once and written once (8 B/pixel) instead of three round trips.  Same arithmetic, same order, same
// untouched borders as k_psf_axial / k_psf_lateral / k_envelope (bit-identical results).
// ------------------------------------------------------------------------------------------------
#define MCRT_FUSED_THREADS 512
#define MCRT_FUSED_MAX_CHUNKS 64     // rows <= 2048 on this path

__device__ __forceinline__ unsigned peak_mask_smem(const float* I, int rows, int c, int lane)
{
    const int i = (c << 5) + lane;
    const float v = i < rows ? I[i] : 0.0f;
    const float vm = (i >= 1 && i < rows) ? I[i - 1] : 0.0f;
    const float vp = (i + 1 < rows) ? I[i + 1] : 0.0f;
    const bool peak = (i >= 1) && (i + 1 < rows) && (vm < v) && !(v < vp);
    return __ballot_sync(0xffffffffu, peak);
}

// KA / KL > 0: compile-time tap counts (taps held in registers, loops fully unrolled); 0: run-time sizes.
template <int KA, int KL>
__global__ void __launch_bounds__(MCRT_FUSED_THREADS) k_post_fused(const float* __restrict__ in, const int cols, const int rows,
                                                                  const float* __restrict__ ax_taps, const int ka,
                                                                  const float* __restrict__ lat_taps, const int kl, const int flags,
                                                                  const int TC, const int col_offset, const int cols_total, float* __restrict__ out,
                                                                  const int in_pitch)
{
    extern __shared__ float sm[];
    __shared__ float s_taps_a[MCRT_MAX_TAPS], s_taps_l[MCRT_MAX_TAPS];
    __shared__ unsigned s_mask[MCRT_FUSED_THREADS / 32][MCRT_FUSED_MAX_CHUNKS];
    __shared__ int s_next[MCRT_FUSED_THREADS / 32][MCRT_FUSED_MAX_CHUNKS];
    const bool conv = (flags & 1) != 0, env = (flags & 2) != 0;
    const int W = conv ? TC + kl - 1 : TC;                 // staged scanlines (tile + right halo)
    float* s_in = sm;                                      // [W][rows]   raw scanlines; the first TC are overwritten
    float* s_ax = sm + (size_t)W * rows;                   // [W][rows]   axial pass result (conv only)
    float* s_out = s_in;                                   //             in place by the lateral pass
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, NW = blockDim.x >> 5;
    const int c0 = blockIdx.x * TC;
    const float* img_in = in + (size_t)blockIdx.y * cols * in_pitch;
    float* img_out = out + (size_t)blockIdx.y * cols * rows;
    if (conv) {
        for (int i = tid; i < ka; i += blockDim.x) s_taps_a[i] = ax_taps[i];
        for (int i = tid; i < kl; i += blockDim.x) s_taps_l[i] = lat_taps[i];
    }
    // stage: a warp walks one scanline, lanes on consecutive rows (coalesced); all loads are independent,
    // so the L2 latency is paid once per warp, not once per tap
    for (int c = w; c < W; c += NW) {
        const int gc = c0 + c;
        float* dst = s_in + (size_t)c * rows;
        if (gc < cols) {
            const float* src = img_in + (size_t)gc * in_pitch;
            for (int r = lane; r < rows; r += 32) dst[r] = __ldg(&src[r]);
        }
    }
    __syncthreads();
    if (conv) {
        // axial pass (rfimage.h:97-108): forward-looking taps, sequential fp32 sum
        for (int c = w; c < W; c += NW) {
            if (c0 + c >= cols) continue;
            const float* src = s_in + (size_t)c * rows;
            float* dst = s_ax + (size_t)c * rows;
            if (KA > 0) {
                float ta[KA > 0 ? KA : 1];
#pragma unroll
                for (int k = 0; k < KA; k++) ta[k] = s_taps_a[k];
                for (int r = KA + lane; r < rows - KA; r += 32) {
                    float convolution = 0;
#pragma unroll
                    for (int k = 0; k < KA; k++) convolution += src[r + k] * ta[k];
                    dst[r] = convolution;
                }
            } else {
                for (int r = ka + lane; r < rows - ka; r += 32) {
                    float convolution = 0;
#pragma unroll 4
                    for (int k = 0; k < ka; k++) convolution += src[r + k] * s_taps_a[k];
                    dst[r] = convolution;
                }
            }
        }
        __syncthreads();
        // lateral pass (rfimage.h:111-122), in place over the raw tile; borders keep the raw samples (B-9)
        for (int c = w; c < TC; c += NW) {
            const int gc = c0 + c;
            if (gc >= cols || !(col_offset + gc >= kl / 2 && col_offset + gc < cols_total - kl)) continue;   // global indices
            float* dst = s_out + (size_t)c * rows;
            if (KL > 0) {
                float tl[KL > 0 ? KL : 1];
#pragma unroll
                for (int k = 0; k < KL; k++) tl[k] = s_taps_l[k];
                for (int r = ka + lane; r < rows - ka; r += 32) {
                    const float* src = s_ax + (size_t)c * rows + r;
                    float convolution = 0;
#pragma unroll
                    for (int k = 0; k < KL; k++) convolution += src[(size_t)k * rows] * tl[k];
                    dst[r] = convolution;
                }
            } else {
                for (int r = ka + lane; r < rows - ka; r += 32) {
                    const float* src = s_ax + (size_t)c * rows + r;
                    float convolution = 0;
#pragma unroll 4
                    for (int k = 0; k < kl; k++) convolution += src[(size_t)k * rows] * s_taps_l[k];
                    dst[r] = convolution;
                }
            }
        }
        __syncthreads();
    }
    if (!env) {
        for (int c = w; c < TC; c += NW) {
            const int gc = c0 + c;
            if (gc >= cols) continue;
            const float* src = s_out + (size_t)c * rows;
            for (int r = lane; r < rows; r += 32) img_out[(size_t)gc * rows + r] = src[r];
        }
        return;
    }
    // envelope (rfimage.h:54-91): one warp per scanline, see k_envelope
    const int n_chunks = (rows + 31) >> 5;
    for (int c = w; c < TC; c += NW) {
        if (c0 + c >= cols) continue;                      // warp-uniform
        const float* I = s_out + (size_t)c * rows;
        float* O = img_out + (size_t)(c0 + c) * rows;
        int next = rows;
        for (int ch = n_chunks - 1; ch >= 0; ch--) {
            const unsigned m = peak_mask_smem(I, rows, ch, lane);
            if (lane == 0) { s_mask[w][ch] = m; s_next[w][ch] = next; }
            if (m) next = (ch << 5) + (__ffs(m) - 1);
        }
        __syncwarp();
        int last_peak = 0;
        for (int ch = 0; ch < n_chunks; ch++) {
            const unsigned mask = s_mask[w][ch];
            const int i = (ch << 5) + lane;
            const unsigned le = mask & (0xffffffffu >> (31 - lane));
            const int p = le ? (ch << 5) + (31 - __clz(le)) : last_peak;
            const unsigned gt = lane == 31 ? 0u : (mask & (0xffffffffu << (lane + 1)));
            const int q = gt ? (ch << 5) + (__ffs(gt) - 1) : s_next[w][ch];
            if (i < rows) {
                float r = I[i];
                if (q < rows) {
                    const float last = (p == 0) ? I[0] : fabsf(I[p]);
                    const float new_peak = fabsf(I[q]);
                    const float alpha = ((float)i - (float)p) / ((float)q - (float)p);
                    r = last * (1 - alpha) + new_peak * alpha;
                }
                O[i] = r;
            }
            if (mask) last_peak = (ch << 5) + (31 - __clz(mask));
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Round 2: the fused post kernel for the reference's image geometry, rebuilt around three facts measured on round 1's
// k_post_fused<7,13> (238 thread instructions per pixel, issue-bound; profiles/r02a_hot_lines.txt):
//  * staging: the raw image arrives with a 16-byte-aligned row pitch (AcqDev::rf_pitch), so the tile + halo scanlines of a
//    CTA are ONE contiguous, aligned span of HBM: a single TMA bulk copy (cp.async.bulk -> UBLKCP) signalled on an mbarrier
//    replaces 12 700 LDG.32 + STS.32 per CTA;
//  * taps: each tap application is one multiply and one add (the reference's un-contracted sum), but the operands now come
//    from registers: the axial pass computes 4 consecutive rows per thread from 3 LDS.128 (was 7 LDS.32 per row), the
//    lateral pass keeps 8 consecutive scanlines of a row pair in registers and streams the 20 scanlines they need past them
//    (2.5 LDS.64 per 2 pixels instead of 13 LDS.32 per pixel); the taps are kernel parameters (constant-bank operands);
//  * the envelope (one warp per scanline) is unchanged.
// Same arithmetic, same order, same untouched borders: bit-identical to k_post_fused and the unfused kernels.
// ------------------------------------------------------------------------------------------------
#define MCRT_TMA_THREADS 1024
#define MCRT_TMA_RUN 8               // scanlines per thread in the lateral pass
#define MCRT_TMA_MAX_CHUNKS 20       // rows <= 640
#define MCRT_TMA_SMEM_LIMIT (200 * 1024)
struct PostTaps { float a[8]; float l[16]; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// a / b for integers 0 <= a < b <= 2048 held in floats, correctly rounded, from the correctly rounded reciprocal y = fl(1 / b):
// Markstein's sequence q0 = a y, r = fma(-q0, b, a) (exact), q = fma(r, y, q0).  Equal to the IEEE quotient for EVERY such pair
// (exhaustive check: tests/test_host_and_numerics.py::test_envelope_alpha_division_is_exact); 3 FP instructions instead of the
// ~10 of the IEEE division the reference's alpha = (j - p) / (q - p) compiles to (rfimage.h:80).
__device__ __forceinline__ float div_small_int(float a, float b, float y)
{
    const float q0 = a * y;
    const float r = __fmaf_rn(-q0, b, a);
    return __fmaf_rn(r, y, q0);
}

// PERSISTENT: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Two shared-memory buffers alternate as
// (raw tile, axial result); while the envelope of tile n runs out of the first, the raw scanlines of tile n + 1 are already
// streaming into the second (its axial values are dead once the lateral pass is done), so no warp ever waits for HBM after
// the first tile.
// NCH > 0: compile-time number of 32-row chunks of a scanline (15 for the reference's 465 rows): the envelope's unrolled chunk
// loops carry no guards; NCH == 0: any rows <= 32 * MCRT_TMA_MAX_CHUNKS.
template <int KA, int KL, int NCH>
__global__ void __launch_bounds__(MCRT_TMA_THREADS, 1) k_post_tma(const float* __restrict__ in, const int cols, const int rows, const int pitch,
                                                                 const __grid_constant__ PostTaps taps, const int TC, const int tiles_per_image,
                                                                 const int n_tiles, const int col_offset, const int cols_total,
                                                                 float* __restrict__ out, const unsigned long long* __restrict__ out_target)
{
    // out_target (nullable): {destination base pointer, image stride in floats} kept in device memory, so that a captured CUDA graph can
    // write every call's frames STRAIGHT into the caller's buffer (and, with a stride, into the interleaved slots of a round-robin sweep)
    // instead of an internal image that is copied afterwards
    float* out_eff = out;
    size_t img_stride = (size_t)cols * rows;
    if (out_target) { out_eff = reinterpret_cast<float*>(__ldg(&out_target[0])); img_stride = (size_t)__ldg(&out_target[1]); }
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ float s_rcp[32 * MCRT_TMA_MAX_CHUNKS + 1];  // fl(1 / b), b = 1 .. rows
    const int W = TC + KL - 1;                             // staged scanlines (tile + right halo)
    float* const buf0 = sm;                                // [W][pitch]
    float* const buf1 = sm + (size_t)W * pitch;            // [W][pitch]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
    // tile t -> first scanline, number of staged scanlines, source address
    auto issue_load = [&](int t, float* dst, unsigned bar) {
        const int img = t / tiles_per_image, c0 = (t - img * tiles_per_image) * TC;
        const int wl = cols - c0 < W ? cols - c0 : W;
        const unsigned bytes = (unsigned)((size_t)wl * pitch * sizeof(float));     // multiple of 16: pitch % 4 == 0
        const float* src = in + ((size_t)img * cols + c0) * pitch;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i <= rows; i += MCRT_TMA_THREADS) s_rcp[i] = i > 0 ? 1.0f / (float)i : 0.0f;
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < n_tiles) issue_load(blockIdx.x, buf0, bar0);
    const int Q = pitch >> 2, P2 = pitch >> 1;
    const int n_chunks = NCH > 0 ? NCH : (rows + 31) >> 5;
    constexpr int kMaxChunks = NCH > 0 ? NCH : MCRT_TMA_MAX_CHUNKS;
    // (scanline, block of 32 quads) items of the axial pass: blocks per scanline is a power of two for the reference geometry
    const int qblocks = (Q + 31) >> 5;
    const int qshift = (qblocks & (qblocks - 1)) == 0 ? 31 - __clz(qblocks) : -1;
    int n = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, n++) {
        float* const s_raw = (n & 1) ? buf1 : buf0;        // raw scanlines; the first TC become the result
        float* const s_ax = (n & 1) ? buf0 : buf1;         // axial pass
        const int img = t / tiles_per_image, c0 = (t - img * tiles_per_image) * TC;
        const int wl = cols - c0 < W ? cols - c0 : W;      // staged scanlines that exist
        float* img_out = out_eff + (size_t)img * img_stride;
        {
            const unsigned bar = (n & 1) ? bar1 : bar0, parity = (unsigned)(n >> 1) & 1u;
            unsigned done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            }
        }
        // ---- axial pass (rfimage.h:97-108): item = (scanline, block of 32 x 4 rows), lane = 4 consecutive rows; forward-looking
        // taps, sequential fp32 sum
        {
            for (int it = w; it < wl * qblocks; it += MCRT_TMA_THREADS / 32) {
                const int c = qshift >= 0 ? it >> qshift : it / qblocks, q = (it - c * qblocks) * 32 + lane;
                if (q >= Q) continue;
                const float4* src4 = reinterpret_cast<const float4*>(s_raw + (size_t)c * pitch);
                float x[12];
                const float4 v0 = src4[q];
                const float4 v1 = q + 1 < Q ? src4[q + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 v2 = q + 2 < Q ? src4[q + 2] : make_float4(0.f, 0.f, 0.f, 0.f);
                x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                x[8] = v2.x; x[9] = v2.y; x[10] = v2.z; x[11] = v2.w;
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float convolution = 0;
#pragma unroll
                    for (int k = 0; k < KA; k++) convolution += x[j + k] * taps.a[k];
                    o[j] = convolution;
                }
                reinterpret_cast<float4*>(s_ax + (size_t)c * pitch)[q] = make_float4(o[0], o[1], o[2], o[3]);   // rows outside [KA, rows - KA) are never read below
            }
        }
        __syncthreads();
        // ---- lateral pass (rfimage.h:111-122), into the raw tile; borders keep the raw samples (B-9).  Thread = one pair of
        // rows x MCRT_TMA_RUN consecutive scanlines; the RUN + KL - 1 axial scanlines they need stream through registers once.
        {
            const int n_runs = TC / MCRT_TMA_RUN;
            const int run = tid / P2, p = tid - run * P2;
            if (run < n_runs) {
                const int cs = run * MCRT_TMA_RUN;
                const float2* ax2 = reinterpret_cast<const float2*>(s_ax) + p;
                float2 acc[MCRT_TMA_RUN];
#pragma unroll
                for (int cc = 0; cc < MCRT_TMA_RUN; cc++) acc[cc] = make_float2(0.f, 0.f);
#pragma unroll
                for (int jj = 0; jj < MCRT_TMA_RUN + KL - 1; jj++) {
                    const int j = cs + jj;
                    const float2 v = j < wl ? ax2[(size_t)j * P2] : make_float2(0.f, 0.f);
#pragma unroll
                    for (int cc = 0; cc < MCRT_TMA_RUN; cc++) {
                        const int k = jj - cc;                      // compile-time after unrolling
                        if (k >= 0 && k < KL) { acc[cc].x += v.x * taps.l[k]; acc[cc].y += v.y * taps.l[k]; }
                    }
                }
                const int r0 = 2 * p;
                const bool ok0 = r0 >= KA && r0 < rows - KA, ok1 = r0 + 1 >= KA && r0 + 1 < rows - KA;
#pragma unroll
                for (int cc = 0; cc < MCRT_TMA_RUN; cc++) {
                    const int c = cs + cc, gc = c0 + c;
                    if (gc < cols && col_offset + gc >= KL / 2 && col_offset + gc < cols_total - KL) {     // global indices
                        float* dst = s_raw + (size_t)c * pitch + r0;
                        if (ok0) dst[0] = acc[cc].x;
                        if (ok1) dst[1] = acc[cc].y;
                    }
                }
            }
        }
        __syncthreads();
        // the axial buffer is dead: start streaming the next tile's raw scanlines into it (generic-proxy accesses to it are
        // ordered before the async-proxy write by the barrier above + the proxy fence)
        if (tid == 0 && t + (int)gridDim.x < n_tiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_load(t + gridDim.x, s_ax, (n & 1) ? bar0 : bar1);
        }
        // ---- envelope (rfimage.h:54-91): one warp per scanline.  A sample i in [1, rows-2] is a peak iff I[i-1] < I[i] and not
        // I[i] < I[i+1]; the peak masks of all 32-row chunks and "first peak after the chunk" stay in (warp-uniform) registers.
        for (int c = w; c < TC; c += MCRT_TMA_THREADS / 32) {
            if (c0 + c >= cols) continue;                      // warp-uniform
            const float* I = s_raw + (size_t)c * pitch;
            float* O = img_out + (size_t)(c0 + c) * rows;
            unsigned mask[kMaxChunks];
            int nxt[kMaxChunks];
            int next = rows;
#pragma unroll
            for (int ch = kMaxChunks - 1; ch >= 0; ch--) {
                mask[ch] = 0u; nxt[ch] = rows;
                if (ch < n_chunks) {
                    const unsigned m = peak_mask_smem(I, rows, ch, lane);
                    mask[ch] = m; nxt[ch] = next;
                    if (m) next = (ch << 5) + (__ffs(m) - 1);
                }
            }
            const float first = I[0];
            int last_peak = 0;
#pragma unroll
            for (int ch = 0; ch < kMaxChunks; ch++) {
                if (ch < n_chunks) {
                    const unsigned m = mask[ch];
                    const int i = (ch << 5) + lane;
                    const unsigned le = m & (0xffffffffu >> (31 - lane));
                    const int p = le ? (ch << 5) + (31 - __clz(le)) : last_peak;
                    const unsigned gt = lane == 31 ? 0u : (m & (0xffffffffu << (lane + 1)));
                    const int q = gt ? (ch << 5) + (__ffs(gt) - 1) : nxt[ch];
                    if (i < rows) {
                        float r = I[i];
                        if (q < rows) {
                            const float last = (p == 0) ? first : fabsf(I[p]);
                            const float new_peak = fabsf(I[q]);
                            const int d = q - p;
                            const float alpha = div_small_int((float)(i - p), (float)d, s_rcp[d]);   // == ((float)i - (float)p) / ((float)q - (float)p)
                            r = last * (1 - alpha) + new_peak * alpha;
                        }
                        O[i] = r;
                    }
                    if (m) last_peak = (ch << 5) + (31 - __clz(m));
                }
            }
        }
        __syncthreads();                                       // the next tile's axial pass overwrites this tile's result buffer
    }
}

// ------------------------------------------------------------------------------------------------
// Log compression: the block the reference keeps commented out in rf_image::postprocess
// (rfimage.h:127-136): max = minMaxLoc(intensities); I = log10(I + 1) / log10(max + 1), applied to the
// envelope image before scan conversion.  Optional (mcrt_set_option "log_compress").
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int f2ord_img(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f_img(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256) k_image_max(const float* __restrict__ img, const int64_t px_per_image, int* __restrict__ max_bits)
{
    const float* I = img + (size_t)blockIdx.y * px_per_image;
    float m = -3.0e38f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px_per_image; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, __ldg(&I[i]));
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = fmaxf(m, s[w]);
        atomicMax(&max_bits[blockIdx.y], f2ord_img(m));
    }
}

__global__ void __launch_bounds__(256) k_log_compress(float* __restrict__ img, const int64_t px_per_image, const int* __restrict__ max_bits)
{
    const double ln10 = 2.30258509299404568402;
    const double maxv = (double)ord2f_img(max_bits[blockIdx.y]);
    const double den = mc_log(maxv + 1) / ln10;                               // std::log10(max + 1), double
    float* I = img + (size_t)blockIdx.y * px_per_image;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px_per_image; i += (int64_t)gridDim.x * blockDim.x) {
        const float num = (float)(mc_log((double)(I[i] + 1)) / ln10);         // std::log10(float)
        I[i] = (float)((double)num / den);
    }
}

__global__ void __launch_bounds__(256) k_copy(const float* __restrict__ in, const int64_t n, float* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

// [n][cols][rows] -> [n][rows][cols] through a padded shared-memory tile
__global__ void k_transpose(const float* __restrict__ in, const int cols, const int rows, float* __restrict__ out)
{
    __shared__ float tile[32][33];
    const size_t img = (size_t)blockIdx.z * cols * rows;
    int r = blockIdx.x * 32 + threadIdx.x, c = blockIdx.y * 32 + threadIdx.y;
    for (int k = 0; k < 32; k += 8)
        if (r < rows && c + k < cols) tile[threadIdx.y + k][threadIdx.x] = in[img + (size_t)(c + k) * rows + r];
    __syncthreads();
    c = blockIdx.y * 32 + threadIdx.x; r = blockIdx.x * 32 + threadIdx.y;
    for (int k = 0; k < 32; k += 8)
        if (c < cols && r + k < rows) out[img + (size_t)(r + k) * cols + c] = tile[threadIdx.x][threadIdx.y + k];
}

// ------------------------------------------------------------------------------------------------
// cv::remap(intensities, scan_converted, map1 = map_y (x = source column), map2 = map_x (y = source
// row), INTER_LINEAR, BORDER_CONSTANT 0) (rfimage.h:139) as OpenCV evaluates it for CV_32FC1 maps:
// coordinates rounded to 1/32 pixel, four taps weighted by a float table, out-of-image taps = 0.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int remap_fix(float v)
{
    const double t = (double)v * 32.0;
    if (!(t == t)) return INT32_MIN;
    if (t <= -2147483648.0) return INT32_MIN;
    if (t >= 2147483647.0) return INT32_MAX;
    return __double2int_rn(t);           // cvRound: round half to even
}

__global__ void __launch_bounds__(256) k_scan_convert(const float* __restrict__ rf, const int n_images, const int cols, const int rows,
                                                     const float* __restrict__ map_x, const float* __restrict__ map_y, const int scan_rows,
                                                     const int scan_cols, float* __restrict__ out)
{
    const int64_t per = (int64_t)scan_rows * scan_cols;
    const int64_t total = per * n_images;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = i / per, px = i - img * per;
        const float* S = rf + img * (int64_t)cols * rows;
        const float x = __ldg(&map_y[px]), y = __ldg(&map_x[px]);
        const int sx = remap_fix(x), sy = remap_fix(y);
        const int ix = sx >> 5, iy = sy >> 5;
        const int fx = sx & 31, fy = sy & 31;
        const float ax = fx * (1.0f / 32), ay = fy * (1.0f / 32);
        const float w0 = (1.0f - ay) * (1.0f - ax), w1 = (1.0f - ay) * ax, w2 = ay * (1.0f - ax), w3 = ay * ax;
        // source pixel (row r, column c) lives at S[c * rows + r] (scanline-major)
        auto pix = [&](int r, int c) -> float { return (r >= 0 && r < rows && c >= 0 && c < cols) ? S[(size_t)c * rows + r] : 0.0f; };
        float v;
        if (ix >= cols || ix + 1 < 0 || iy >= rows || iy + 1 < 0) v = 0.0f;
        else v = pix(iy, ix) * w0 + pix(iy, ix + 1) * w1 + pix(iy + 1, ix) * w2 + pix(iy + 1, ix + 1) * w3;
        out[i] = v;
    }
}

int grid1d(int64_t n, int block)
{
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

size_t accumulate_columns_bytes(const AcqDev& aq, int n_poses)
{
    return sizeof(float) * (size_t)n_poses * aq.elements * aq.samples * aq.rows;
}

bool accumulate_windowed_supported(const SceneDev& sc, const AcqDev& aq)
{
    return aq.samples >= 1 && aq.samples <= 128 && sc.spacing[0] == 1.0f && sc.spacing[1] == 1.0f && sc.spacing[2] == 1.0f;
}

cudaError_t launch_accumulate(const SceneDev& sc, const AcqDev& aq, const float2* d_volume, const DevSegment* d_segments,
                              const int32_t* d_nseg, int n_poses, float* d_rf, unsigned long long* d_steps, float* d_columns,
                              cudaStream_t stream, int* launches)
{
    const int n_paths = n_poses * aq.elements * aq.samples;
    if (aq.accumulate_windowed && accumulate_windowed_supported(sc, aq)) {
        // row-window synchronous kernel: no columns in HBM, sample reduction inside (d_steps[1] counts late echoes)
        const bool warp = aq.samples <= 32;                       // a warp owns whole scanlines: __syncwarp instead of a CTA barrier
        const int G = (warp ? 32 : 128) / aq.samples;
        const int n_scanlines = n_poses * aq.elements;
        const int n_groups = (n_scanlines + G - 1) / G;
        const size_t smem = sizeof(float) * MCRT_WIN_RING * (size_t)win_stride(G * aq.samples) * (warp ? 4 : 1);
        const int grid = warp ? (n_groups + 3) / 4 : n_groups;
        if (warp) {
            if (aq.samples == 16) {            // BASELINE's 16 samples per element: compile-time sample count
                if (aq.voxel_fma_division)
                    k_accumulate_win<true, true, 16><<<grid, 128, smem, stream>>>(sc, aq, d_volume, d_segments, d_nseg, n_scanlines, G, d_rf, d_steps, d_steps + 1);
                else
                    k_accumulate_win<false, true, 16><<<grid, 128, smem, stream>>>(sc, aq, d_volume, d_segments, d_nseg, n_scanlines, G, d_rf, d_steps, d_steps + 1);
            } else if (aq.voxel_fma_division)
                k_accumulate_win<true, true, 0><<<grid, 128, smem, stream>>>(sc, aq, d_volume, d_segments, d_nseg, n_scanlines, G, d_rf, d_steps, d_steps + 1);
            else
                k_accumulate_win<false, true, 0><<<grid, 128, smem, stream>>>(sc, aq, d_volume, d_segments, d_nseg, n_scanlines, G, d_rf, d_steps, d_steps + 1);
        } else {
            if (aq.voxel_fma_division)
                k_accumulate_win<true, false, 0><<<grid, 128, smem, stream>>>(sc, aq, d_volume, d_segments, d_nseg, n_scanlines, G, d_rf, d_steps, d_steps + 1);
            else
                k_accumulate_win<false, false, 0><<<grid, 128, smem, stream>>>(sc, aq, d_volume, d_segments, d_nseg, n_scanlines, G, d_rf, d_steps, d_steps + 1);
        }
        if (launches) (*launches) += 1;
        return cudaGetLastError();
    }
    if (!d_columns) return cudaErrorInvalidValue;
    const int block = 128;
    if (aq.voxel_fma_division)
        k_accumulate<true><<<(n_paths + block - 1) / block, block, 0, stream>>>(sc, aq, d_volume, d_segments, d_nseg, n_paths, d_columns, d_steps);
    else
        k_accumulate<false><<<(n_paths + block - 1) / block, block, 0, stream>>>(sc, aq, d_volume, d_segments, d_nseg, n_paths, d_columns, d_steps);
    const int64_t n_pixels = (int64_t)n_poses * aq.elements * aq.rows;
    k_reduce_samples<<<grid1d(n_pixels, 256), 256, 0, stream>>>(d_columns, n_pixels, aq.samples, d_rf, aq.rows, aq.rf_pitch);
    if (launches) (*launches) += 2;
    return cudaGetLastError();
}

cudaError_t launch_accumulate_tree(const SceneDev& sc, const AcqDev& aq, const float2* d_volume, const TreeBuffers& tb, int n_poses,
                                   float* d_rf, unsigned long long* d_steps, cudaStream_t stream, int* launches)
{
    // (no spacing restriction here: time is monotone inside ONE segment whatever the spacing; the single-path kernel needs it across
    // the consecutive segments of a path)
    const int n_scanlines = n_poses * aq.elements;
    cudaError_t e = cudaMemsetAsync(d_rf, 0, sizeof(float) * (size_t)n_scanlines * aq.rf_pitch, stream);
    if (e != cudaSuccess) return e;
    const size_t smem = sizeof(float) * MCRT_WIN_RING * (size_t)win_stride(32) * 4;
    const int grid = (n_scanlines + 3) / 4;                       // a warp per scanline
    if (aq.voxel_fma_division)
        k_accumulate_win<true, true, 32, true><<<grid, 128, smem, stream>>>(sc, aq, d_volume, tb.segments, nullptr, n_scanlines, 1, d_rf, d_steps, d_steps + 1,
                                                                            tb.level_first, tb.level_end, aq.max_depth, tb.n_scanlines);
    else
        k_accumulate_win<false, true, 32, true><<<grid, 128, smem, stream>>>(sc, aq, d_volume, tb.segments, nullptr, n_scanlines, 1, d_rf, d_steps, d_steps + 1,
                                                                             tb.level_first, tb.level_end, aq.max_depth, tb.n_scanlines);
    if (launches) (*launches) += 1;
    return cudaGetLastError();
}

// shared-memory budget of the fused path
#define MCRT_FUSED_SMEM_LIMIT (100 * 1024)     // + ~12 KB static: two CTAs of 512 threads per SM
static size_t fused_smem_bytes(int rows, int kl, int flags, int tc)
{
    return sizeof(float) * (size_t)rows * ((flags & 1) ? (size_t)(2 * (tc + kl - 1)) : (size_t)tc);
}
// widest tile (<= 32 scanlines, >= 4) whose staging fits; 0 = use the unfused kernels
static int fused_tile_cols(int rows, int kl, int flags, size_t* smem)
{
    if (rows > 32 * MCRT_FUSED_MAX_CHUNKS) return 0;
    for (int tc = 32; tc >= 4; tc--) {
        const size_t bytes = fused_smem_bytes(rows, kl, flags, tc);
        if (bytes <= MCRT_FUSED_SMEM_LIMIT) { *smem = bytes; return tc; }
    }
    return 0;
}

cudaError_t validate_fma_division(float resolution, bool* ok)
{
    unsigned int* d_bad = nullptr;
    unsigned int bad = 1;
    cudaError_t e = cudaMalloc(&d_bad, sizeof(unsigned int));
    if (e != cudaSuccess) return e;
    cudaMemset(d_bad, 0, sizeof(unsigned int));
    k_validate_fma_division<<<(1u << 24) / 256, 256>>>(resolution, d_bad);       // 2^24 threads x 256 bit patterns
    e = cudaMemcpy(&bad, d_bad, sizeof(unsigned int), cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    *ok = (e == cudaSuccess) && bad == 0 && resolution > 0.0f;
    return e;
}

cudaError_t init_image_kernels()
{
    // per-device function attribute; must not be issued inside a stream capture
    cudaError_t e = cudaFuncSetAttribute(k_post_fused<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MCRT_FUSED_SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_envelope_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);     // rows <= 32768
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_post_tma<7, 13, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, MCRT_TMA_SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_post_tma<7, 13, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, MCRT_TMA_SMEM_LIMIT);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_post_fused<7, 13>, cudaFuncAttributeMaxDynamicSharedMemorySize, MCRT_FUSED_SMEM_LIMIT);
}

int post_launch_count(int cols, int rows, int n_lateral, int flags, int n_images)
{
    size_t smem = 0;
    (void)cols;
    if ((flags & 3) && fused_tile_cols(rows, n_lateral, flags, &smem) > 0 && n_images <= 65535) return 1;
    const int64_t n_scanlines = (int64_t)n_images * cols;
    return ((flags & 1) ? (int)((n_scanlines + 65534) / 65535) + 1 : 0) + ((flags & 2) ? 1 : ((flags & 1) ? 0 : 1));
}

// the TMA-staged kernel needs both passes, the reference's tap counts, a 16-byte row pitch and scanlines that fit shared memory
static bool post_tma_usable(int rows, int pitch, int n_axial, int n_lateral, int flags, const float* h_axial, const float* h_lateral)
{
    return flags == 3 && n_axial == 7 && n_lateral == 13 && (pitch & 3) == 0 && pitch >= rows && rows <= 32 * MCRT_TMA_MAX_CHUNKS && rows > 2 * 7 &&
           h_axial && h_lateral && pitch * 2 <= MCRT_TMA_THREADS;
}
static size_t post_tma_smem(int pitch, int tc) { return sizeof(float) * 2 * (size_t)(tc + 13 - 1) * pitch; }
// A/B switch (option "long_ct"): compile-time-tap axial / lateral kernels and the shared-memory envelope on the long-scanline path
static bool g_long_ct = true;
void set_long_scanline_ct(bool on) { g_long_ct = on; }

// scanlines per tile of the TMA-staged kernel for this call, 0 when another kernel has to take it
static int post_tma_tile_cols(int n_images, int cols, int rows, int in_pitch, int n_axial, int n_lateral, int flags, const float* h_axial,
                              const float* h_lateral, bool by_row)
{
    if (by_row || !((int64_t)n_images * cols < 0x7fffffff) || !post_tma_usable(rows, in_pitch, n_axial, n_lateral, flags, h_axial, h_lateral)) return 0;
    // widest tile of 8 / 16 / 32 scanlines that fits; few images (latency mode): narrower tiles so the grid still covers the SMs
    int tc = 32;
    while (tc > 8 && (post_tma_smem(in_pitch, tc) > MCRT_TMA_SMEM_LIMIT || (int64_t)((cols + tc - 1) / tc) * n_images < 148)) tc >>= 1;
    return (post_tma_smem(in_pitch, tc) <= MCRT_TMA_SMEM_LIMIT && (in_pitch / 2) * (tc / MCRT_TMA_RUN) <= MCRT_TMA_THREADS) ? tc : 0;
}

bool post_writes_through_target(int n_images, int cols, int rows, int in_pitch, int n_axial, int n_lateral, int flags, const float* h_axial,
                                const float* h_lateral, bool by_row)
{
    if (in_pitch <= 0) in_pitch = rows;
    return post_tma_tile_cols(n_images, cols, rows, in_pitch, n_axial, n_lateral, flags, h_axial, h_lateral, by_row) > 0;
}

void launch_post(const float* d_in, int n_images, int cols, int rows, const float* d_axial, int n_axial, const float* d_lateral,
                 int n_lateral, int flags, float* d_tmp0, float* d_tmp1, float* d_out, cudaStream_t stream, int* launches, int col_offset,
                 int cols_total, const float* d_lateral_by_row, int in_pitch, const float* h_axial, const float* h_lateral,
                 const unsigned long long* d_out_target)
{
    if (cols_total <= 0) { col_offset = 0; cols_total = cols; }
    if (in_pitch <= 0) in_pitch = rows;
    {
        const int tc = post_tma_tile_cols(n_images, cols, rows, in_pitch, n_axial, n_lateral, flags, h_axial, h_lateral, d_lateral_by_row != nullptr);
        if (tc > 0) {
            PostTaps taps;
            for (int k = 0; k < 8; k++) taps.a[k] = k < 7 ? h_axial[k] : 0.0f;
            for (int k = 0; k < 16; k++) taps.l[k] = k < 13 ? h_lateral[k] : 0.0f;
            const int tiles_per_image = (cols + tc - 1) / tc;
            const int64_t n_tiles = (int64_t)tiles_per_image * n_images;
            int sms = 148;
            { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
            const int grid = (int)(n_tiles < sms ? n_tiles : sms);                 // persistent: one CTA per SM
            if (((rows + 31) >> 5) == 15)          // the reference's 465-row image (main.cpp:30, rfimage.h:180)
                k_post_tma<7, 13, 15><<<grid, MCRT_TMA_THREADS, post_tma_smem(in_pitch, tc), stream>>>(d_in, cols, rows, in_pitch, taps, tc, tiles_per_image,
                                                                                                     (int)n_tiles, col_offset, cols_total, d_out, d_out_target);
            else
                k_post_tma<7, 13, 0><<<grid, MCRT_TMA_THREADS, post_tma_smem(in_pitch, tc), stream>>>(d_in, cols, rows, in_pitch, taps, tc, tiles_per_image,
                                                                                                    (int)n_tiles, col_offset, cols_total, d_out, d_out_target);
            if (launches) (*launches)++;
            return;
        }
    }
    const int64_t n_scanlines = (int64_t)n_images * cols;
    const int64_t total = n_scanlines * rows;
    size_t smem = 0;
    int tc = (flags & 3) ? fused_tile_cols(rows, n_lateral, flags, &smem) : 0;
    // few images (latency mode): narrower tiles so the grid still covers the 148 SMs
    while (tc > 4 && (int64_t)((cols + tc - 1) / tc) * n_images < 2 * 148) {
        tc = tc > 8 ? tc / 2 : 4;
        smem = fused_smem_bytes(rows, n_lateral, flags, tc);
    }
    if (tc > 0 && n_images <= 65535 && !d_lateral_by_row) {       // the fused kernel holds ONE set of lateral taps in registers
        // whole scanlines fit shared memory: one pass over HBM
        dim3 grid((cols + tc - 1) / tc, n_images, 1);
        if (n_axial == 7 && n_lateral == 13)      // the reference's psf<7,13,...> (main.cpp:34)
            k_post_fused<7, 13><<<grid, MCRT_FUSED_THREADS, smem, stream>>>(d_in, cols, rows, d_axial, 7, d_lateral, 13, flags, tc, col_offset, cols_total, d_out, in_pitch);
        else
            k_post_fused<0, 0><<<grid, MCRT_FUSED_THREADS, smem, stream>>>(d_in, cols, rows, d_axial, n_axial, d_lateral, n_lateral, flags, tc, col_offset, cols_total, d_out, in_pitch);
        if (launches) (*launches)++;
        return;
    }
    // long scanlines: register-blocked axial / lateral passes, mask-based parallel envelope
    const float* cur = d_in;
    // compile-time tap loops (taps as kernel parameters) whenever the host copy of the taps is at hand and they fit
    LongTaps ltaps_a, ltaps_l;
    const int ng_ax = (n_axial + MCRT_PSF_RA - 1) / MCRT_PSF_RA, ng_lat = (n_lateral + MCRT_PSF_RL - 1) / MCRT_PSF_RL;
    const bool ct_ax = h_axial && n_axial >= 1 && n_axial <= 64 && g_long_ct;
    const bool ct_lat = h_lateral && n_lateral >= 1 && n_lateral <= 64 && !d_lateral_by_row && g_long_ct;
    for (int k = 0; k < 64; k++) { ltaps_a.t[k] = ct_ax && k < n_axial ? h_axial[k] : 0.0f; ltaps_l.t[k] = ct_lat && k < n_lateral ? h_lateral[k] : 0.0f; }
    if (flags & 1) {
        // even chunks: as few CTAs per scanline as fit MCRT_PSF_CHUNK, all the same size
        const int n_ch = (rows + MCRT_PSF_CHUNK - 1) / MCRT_PSF_CHUNK;
        const int chunk_rows = (((rows + n_ch - 1) / n_ch) + MCRT_PSF_RA - 1) / MCRT_PSF_RA * MCRT_PSF_RA;
        const size_t smem_ax = sizeof(float) * (size_t)(psf_pad(chunk_rows + n_axial + 2 * MCRT_PSF_RA) + 1);
        dim3 ga((rows + chunk_rows - 1) / chunk_rows, (unsigned)n_scanlines, 1);
        if (n_scanlines > 65535) ga = dim3(ga.x, 65535, 1);   // guarded below
        for (int64_t s0 = 0; s0 < n_scanlines; s0 += 65535) {
            const int64_t ns = n_scanlines - s0 < 65535 ? n_scanlines - s0 : 65535;
            ga.y = (unsigned)ns;
            if (ct_ax) {
                const size_t smem_ct = sizeof(float) * (size_t)(psf_pad(chunk_rows + ng_ax * MCRT_PSF_RA + 2 * MCRT_PSF_RA) + 1);
                const float* a_in = cur + s0 * in_pitch;
                float* a_out = d_tmp0 + s0 * rows;
#define MCRT_AX_CASE(G) case G: k_psf_axial_ct<G><<<ga, 256, smem_ct, stream>>>(a_in, rows, ltaps_a, n_axial, a_out, in_pitch, chunk_rows); break;
                switch (ng_ax) { MCRT_AX_CASE(1) MCRT_AX_CASE(2) MCRT_AX_CASE(3) MCRT_AX_CASE(4) MCRT_AX_CASE(5) MCRT_AX_CASE(6) MCRT_AX_CASE(7) MCRT_AX_CASE(8) default: break; }
#undef MCRT_AX_CASE
            } else
                k_psf_axial<<<ga, 256, smem_ax, stream>>>(cur + s0 * in_pitch, rows, d_axial, n_axial, d_tmp0 + s0 * rows, in_pitch, chunk_rows);
            if (launches) (*launches)++;
        }
        float* dst = (flags & 2) ? d_tmp1 : d_out;
        dim3 gl((rows + 255) / 256, (cols + MCRT_PSF_R - 1) / MCRT_PSF_R, n_images);
        if (d_lateral_by_row)
            k_psf_lateral<true><<<gl, 256, 0, stream>>>(cur, d_tmp0, cols, rows, d_lateral, n_axial, n_lateral, col_offset, cols_total, dst, d_lateral_by_row, in_pitch);
        else if (ct_lat) {
            const dim3 gc((rows + 255) / 256, (cols + MCRT_PSF_RL - 1) / MCRT_PSF_RL, n_images);
#define MCRT_LAT_CASE(G) case G: k_psf_lateral_ct<G><<<gc, 256, 0, stream>>>(cur, d_tmp0, cols, rows, ltaps_l, n_axial, n_lateral, col_offset, cols_total, dst, in_pitch); break;
            switch (ng_lat) { MCRT_LAT_CASE(1) MCRT_LAT_CASE(2) MCRT_LAT_CASE(3) MCRT_LAT_CASE(4) default: break; }
#undef MCRT_LAT_CASE
        } else
            k_psf_lateral<false><<<gl, 256, 0, stream>>>(cur, d_tmp0, cols, rows, d_lateral, n_axial, n_lateral, col_offset, cols_total, dst, nullptr, in_pitch);
        cur = dst;
        if (launches) (*launches)++;
    }
    if (flags & 2) {
        // one CTA per scanline; the per-chunk peak masks live in shared memory (3 words per 32 rows)
        const int words = (rows + 31) >> 5;
        const size_t smem_env = sizeof(unsigned) * 3 * (size_t)words;
        int64_t g = n_scanlines;
        if (g > 148 * 64) g = 148 * 64;
        const int nw_t = (words + MCRT_ENVT_CH - 1) / MCRT_ENVT_CH;           // warps (tiles) per scanline: <= 32 for rows <= 32768
        const size_t smem_t = sizeof(unsigned) * ((size_t)(rows < MCRT_ENV_RCP ? rows : MCRT_ENV_RCP) + 1 + (size_t)nw_t * (7 * MCRT_ENVT_CH) + 2 * (size_t)nw_t * 4);
        if (g_long_ct && nw_t <= 32) {                                         // (MCRT_ENVT_CH = 32: every rows <= 32768)
            // persistent: a few CTAs per SM walk the scanlines (the reciprocal table is built once per CTA)
            int64_t gt = n_scanlines;
            const int64_t cap = 148 * (int64_t)(2048 / (32 * nw_t) > 0 ? 2048 / (32 * nw_t) : 1) * 2;
            if (gt > cap) gt = cap;
            if (nw_t <= 12) k_envelope_tiles<384><<<(int)gt, 32 * nw_t, smem_t, stream>>>(cur, n_scanlines, rows, d_out);
            else if (nw_t <= 20) k_envelope_tiles<640><<<(int)gt, 32 * nw_t, smem_t, stream>>>(cur, n_scanlines, rows, d_out);
            else k_envelope_tiles<1024><<<(int)gt, 32 * nw_t, smem_t, stream>>>(cur, n_scanlines, rows, d_out);
        } else
            k_envelope_stream<<<(int)g, MCRT_ENV_WARPS * 32, smem_env, stream>>>(cur, n_scanlines, rows, words, d_out);
        if (launches) (*launches) += 1;
    } else if (!(flags & 1)) {
        k_copy<<<grid1d(total, 256), 256, 0, stream>>>(cur, total, d_out);
        if (launches) (*launches)++;
    }
}

int post_preferred_pitch(int rows, int n_axial, int n_lateral)
{
    const int pitch = (rows + 3) & ~3;
    const float dummy = 0.0f;
    return post_tma_usable(rows, pitch, n_axial, n_lateral, 3, &dummy, &dummy) && post_tma_smem(pitch, 8) <= MCRT_TMA_SMEM_LIMIT ? pitch : rows;
}

// ------------------------------------------------------------------------------------------------
// Elevational PSF: the third separable pass.  Every output frame is traced as n_planes ray fans offset along the elevation
// axis; their raw RF images are combined with the elevation taps before the axial / lateral passes (the passes are linear, so
// the order does not matter mathematically; combining first costs one pass over the raw images):
//   out[i] = sum_j in[plane j][i] * w[j], j ascending, separate multiply and add (the oracle's order).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_elevation_combine(const float* __restrict__ in, const int64_t n_out_px, const int64_t px_per_image,
                                                          const int n_planes, const float* __restrict__ w, float* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out_px; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = i / px_per_image, px = i - img * px_per_image;
        const float* src = in + img * n_planes * px_per_image + px;
        float acc = 0.0f;
        for (int j = 0; j < n_planes; j++) acc += __ldg(&src[(int64_t)j * px_per_image]) * __ldg(&w[j]);
        out[i] = acc;
    }
}

void launch_elevation_combine(const float* d_in, int n_out_images, int64_t px_per_image, int n_planes, const float* d_w, float* d_out,
                              cudaStream_t stream, int* launches)
{
    const int64_t n = (int64_t)n_out_images * px_per_image;
    k_elevation_combine<<<grid1d(n, 256), 256, 0, stream>>>(d_in, n, px_per_image, n_planes, d_w, d_out);
    if (launches) (*launches)++;
}

void launch_log_compress(float* d_img, int n_images, int64_t px_per_image, int* d_max_bits, cudaStream_t stream, int* launches)
{
    // ordered-int encoding of -FLT_MAX-ish: any finite sample is larger
    cudaMemsetAsync(d_max_bits, 0x80, sizeof(int) * (size_t)n_images, stream);
    int gx = (int)((px_per_image + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    if (gx > 1024) gx = 1024;
    for (int i0 = 0; i0 < n_images; i0 += 65535) {
        const int ni = n_images - i0 < 65535 ? n_images - i0 : 65535;
        dim3 grid(gx, ni, 1);
        k_image_max<<<grid, 256, 0, stream>>>(d_img + (size_t)i0 * px_per_image, px_per_image, d_max_bits + i0);
        k_log_compress<<<grid, 256, 0, stream>>>(d_img + (size_t)i0 * px_per_image, px_per_image, d_max_bits + i0);
        if (launches) (*launches) += 2;
    }
}

// ------------------------------------------------------------------------------------------------
// B-mode display chain (the reference stops at the envelope and keeps its log compression
// commented out, rfimage.h:127-136): time-gain compensation, log compression to a dynamic range, 8-bit output.
//   v = |E[row]| * gain[row]          gain[row] = 10^((gain_db + tgc_db_per_cm * depth_cm(row)) / 20)   (host table)
//   y = clamp(1 + (20 / DR) * log10(v / max_image(v)), 0, 1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bmode_gain_max(const float* __restrict__ env, const int rows, const int64_t px_per_image,
                                                       const float* __restrict__ gain, float* __restrict__ out, int* __restrict__ max_bits)
{
    const float* I = env + (size_t)blockIdx.y * px_per_image;
    float* O = out + (size_t)blockIdx.y * px_per_image;
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px_per_image; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = fabsf(__ldg(&I[i])) * __ldg(&gain[(int)(i % rows)]);
        O[i] = v;
        m = fmaxf(m, v);                                                    // NaN samples never win
    }
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = fmaxf(m, s[w]);
        atomicMax(&max_bits[blockIdx.y], __float_as_int(m));                // m >= 0: the bit pattern is ordered
    }
}

__global__ void __launch_bounds__(256) k_bmode_compress(float* __restrict__ img, const int64_t px_per_image, const int* __restrict__ max_bits,
                                                       const float dynamic_range_db)
{
    const double ln10 = 2.30258509299404568402;
    const double maxv = (double)__int_as_float(max_bits[blockIdx.y]);
    const double scale = 20.0 / (double)dynamic_range_db;
    float* I = img + (size_t)blockIdx.y * px_per_image;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px_per_image; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = I[i];
        float y = 0.0f;
        if (maxv > 0.0 && v > 0.0f) {
            y = (float)(1.0 + scale * (mc_log((double)v / maxv) / ln10));
            y = y < 0.0f ? 0.0f : (y > 1.0f ? 1.0f : y);
        }
        I[i] = y;                                                            // NaN in -> 0
    }
}

__global__ void __launch_bounds__(256) k_quantize8(const float* __restrict__ in, const int64_t n, unsigned char* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = in[i] * 255.0f;                                       // convertTo(CV_8U, 255.0): round to nearest even, saturate
        const int q = (v != v) ? 0 : __float2int_rn(fminf(fmaxf(v, 0.0f), 255.0f));
        out[i] = (unsigned char)q;
    }
}

void launch_bmode(const float* d_env, int n_images, int cols, int rows, const float* d_gain, float dynamic_range_db, float* d_out,
                  int* d_max_bits, cudaStream_t stream, int* launches)
{
    const int64_t px = (int64_t)cols * rows;
    cudaMemsetAsync(d_max_bits, 0, sizeof(int) * (size_t)n_images, stream);
    int gx = (int)((px + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    if (gx > 1024) gx = 1024;
    for (int i0 = 0; i0 < n_images; i0 += 65535) {
        const int ni = n_images - i0 < 65535 ? n_images - i0 : 65535;
        dim3 grid(gx, ni, 1);
        k_bmode_gain_max<<<grid, 256, 0, stream>>>(d_env + (size_t)i0 * px, rows, px, d_gain, d_out + (size_t)i0 * px, d_max_bits + i0);
        k_bmode_compress<<<grid, 256, 0, stream>>>(d_out + (size_t)i0 * px, px, d_max_bits + i0, dynamic_range_db);
        if (launches) (*launches) += 2;
    }
}

void launch_quantize8(const float* d_in, int64_t n, unsigned char* d_out, cudaStream_t stream, int* launches)
{
    k_quantize8<<<grid1d(n, 256), 256, 0, stream>>>(d_in, n, d_out);
    if (launches) (*launches)++;
}

void launch_transpose(const float* d_in, int n_images, int cols, int rows, float* d_out, cudaStream_t stream, int* launches)
{
    dim3 block(32, 8, 1), grid((rows + 31) / 32, (cols + 31) / 32, n_images);
    k_transpose<<<grid, block, 0, stream>>>(d_in, cols, rows, d_out);
    if (launches) (*launches)++;
}

void launch_scan_convert(const float* d_rf, int n_images, int cols, int rows, const float* d_map_x, const float* d_map_y, int scan_rows,
                         int scan_cols, float* d_out, cudaStream_t stream, int* launches)
{
    const int64_t total = (int64_t)n_images * scan_rows * scan_cols;
    k_scan_convert<<<grid1d(total, 256), 256, 0, stream>>>(d_rf, n_images, cols, rows, d_map_x, d_map_y, scan_rows, scan_cols, d_out);
    if (launches) (*launches)++;
}

}  // namespace mcrt
