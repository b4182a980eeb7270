// align_fill_short.cuh -- K1: the short-pair DP fill (BASELINE config 2: 1M x 150x150).
//
// One THREAD owns two whole pairs, one in each 16-bit half of its registers, so a warp works on
// 64 pairs with no shuffles, no idle lanes and no wavefront fill/drain. Rows are processed in
// blocks of 32 (registers), columns are swept left to right; the bottom row of a block is parked
// in an L1/L2-resident scratch row and picked up by the next block.
//
// Arithmetic (bit-exact restatement of team_alignment.cpp:104-114, ties diagonal > left > up):
// every candidate is carried as 4*score + tag with tag 2 = diagonal, 1 = left, 0 = up, so a plain
// signed max implements both the maximum and the reference's tie order, and the low two bits of
// the winner ARE the direction (stored as code = 2 - tag... see kTagToCode). Per pair of cells:
//     S   = PRMT(tabA, tabB, sel[r])              substitution term from a per-column byte table
//     m1  = VIADDMNMX.S16x2(diag, S, left)        max(diag + S, left)
//     Z   = VIADDMNMX.S16x2(up, 4*gap - 1, m1)    max(up + 4*gap - 1, m1)
//     Y   = LOP3 (Z & 0xFFFCFFFC) | 0x00010001    strip the tag, re-arm the 'left' tag
//     accZ = accZ*4 + Z ; accY = accY*4 + Y       IMAD chains; accZ - accY + 0x5555.. = 8 tags / half
// = 3 alu-pipe + 2 fma-pipe + 1 PRMT issue slots for TWO cells (see the moving-frame note below).
//
// All three alignment types: global captures cell (Q,T); semiGlobal tracks the best of the last column
// (per block, at the pair's last column) and of row Q (every column of the pair's last block), with the
// reference's tie rules (team_alignment.cpp:265-278); local clamps at 0 (tag 3 = stop) and keeps the first
// maximum in row-major order (:186-192): a packed max tree per column gives the column maximum, and only
// when it beats -- or, inside the same block, ties -- the running best are the rows scanned for the cell.
//
// Eligibility (decided on the host, align_plan.cu): both sequences pure ACGT, |s - gap| <= 31 for
// s in {match, mismatch}, and 4 * ((Q+T+2) * max|score| + |gap| * T + 4) <= 32767 so nothing leaves int16.
#pragma once
#include "common.cuh"

namespace b200 {


// Sequences packed 2 bits per base, 16 bases per word, base k of a word at bits [2k, 2k+1].
// Codes: A=0 C=1 T=2 G=3. Pair p's query words start at PairDesc::qpk_off (see pack_kernel).
__device__ __forceinline__ uint32_t acgt_code(uint32_t c) { return (c >> 1) & 3u; }  // A=0 C=1 T=2 G=3 on ASCII
// (ASCII: A=0x41 -> 0, C=0x43 -> 1, G=0x47 -> 3, T=0x54 -> 2; any bijection works for equality.)

// One thread per packed word: classifies its 16 bases (flags, only touched when something other
// than ACGT shows up, so flags[] must be zeroed first) and writes the 2-bit copy. blockIdx.y picks
// query (0) or target (1); `wpp` = words per sequence of the longest short pair.
__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t* __restrict__ qbuf, const uint8_t* __restrict__ tbuf,
            const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work,
            uint32_t wpp, uint8_t* __restrict__ flags, uint32_t* __restrict__ qpk,
            uint32_t* __restrict__ tpk, uint32_t* __restrict__ n_flagged) {
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t wi = (uint32_t)(id / wpp), w = (uint32_t)(id % wpp);
    if (wi >= n_work) return;
    const uint32_t p = work[wi];
    const PairDesc pd = pairs[p];
    const bool which = blockIdx.y != 0;
    const uint32_t len = which ? pd.T : pd.Q;
    if (w * 16 >= len) return;
    const uint8_t* s = (which ? tbuf + pd.t_off : qbuf + pd.q_off) + w * 16;
    const uint32_t nb = min(16u, len - w * 16);
    // 16 bases = up to five aligned 32-bit words; realign with funnel shifts (the buffers are padded, so
    // reading the aligned words around the sequence is safe), then 4 bases per word in parallel.
    const uintptr_t addr = reinterpret_cast<uintptr_t>(s);
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
    const uint32_t shb = (uint32_t)(addr & 3u) * 8u;
    uint32_t raw[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) raw[q] = ((uint32_t)q * 4u < (uint32_t)(addr & 3u) + nb) ? __ldg(aw + q) : 0x41414141u;
    uint32_t word = 0;
    bool dash = false, other = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t v = __funnelshift_r(raw[q], raw[q + 1], shb);      // bases 4q .. 4q+3, base 4q in the low byte
        const uint32_t valid = nb > 4u * q ? min(4u, nb - 4u * q) : 0u;
        if (valid < 4u) v = (valid == 0u) ? 0x41414141u : ((v & (0xffffffffu >> (8u * (4u - valid)))) | (0x41414141u << (8u * valid)));
        const uint32_t c4 = (v >> 1) & 0x03030303u;                 // A=0 C=1 T=2 G=3 per byte
        // valid iff re-encoding the codes gives the bytes back: letters[code] with letters = "ACTG"
        const uint32_t sel = (c4 & 0x3u) | ((c4 >> 4) & 0x30u) | ((c4 >> 8) & 0x300u) | ((c4 >> 12) & 0x3000u);   // one nibble per base
        const uint32_t back = __byte_perm(0x47544341u, 0u, sel);
        if (back != v) {
            other = true;
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) dash |= (((v >> (8 * bb)) & 0xffu) == (uint32_t)'-');
        }
        const uint32_t c8 = (c4 | (c4 >> 6) | (c4 >> 12) | (c4 >> 18)) & 0xffu;     // four 2-bit codes
        word |= c8 << (8 * q);
    }
    (which ? tpk + pd.tpk_off : qpk + pd.qpk_off)[w] = word;
    if (dash || other) {
        const uint32_t f = (dash ? kFlagDash : 0u) | (other ? kFlagNonACGT : 0u);
        uint32_t* word32 = reinterpret_cast<uint32_t*>(flags + (p & ~3u));
        const uint32_t old = atomicOr(word32, f << (8 * (p & 3u)));
        if (((old >> (8 * (p & 3u))) & 0xffu) == 0) atomicAdd(n_flagged, 1u);
    }
}

struct ShortConsts {
    uint32_t tab_diff;    // (S'match ^ S'mismatch) in byte 0
    uint32_t tab_mis;     // S'mismatch in all four bytes
    uint32_t cu;          // 4*gap - 1 in both halves: what the 'up' candidate adds
    // Passed as parameters (constant bank) rather than literals on purpose: LOP3 can take only one
    // immediate, and a literal multiplier of 4 would be strength-reduced onto the (saturated) alu pipe.
    uint32_t mask, one, four;
    int gap, init;
    // substitution terms as int16 (both halves of a table word are built from them) and an opaque 1 for the
    // address multiply-add of the shared-memory lookup (see SubstTable)
    int sm, sx;
    uint32_t one32;
};

// Substitution term from shared memory instead of PRMT (PRMT issues at half rate on the saturated alu pipe: 2 of
// the 5 alu slots a cell pair costs; the LSU pipe is idle). A table word holds both halves' terms; it is indexed
// by the two row codes (a per-register constant) PLUS the two column codes (one value per column): with
// ia = qa + ((4 - ca) & 3) in 0..6, equality of the 2-bit codes is (ia & 3) == 0, so
//     index = ia | ib << 3        = (qa | qb << 3) + (((4 - ca) & 3) | ((4 - cb) & 3) << 3)
// is a plain sum of a row part and a column part (one IMAD on the fma pipe) and the table has 64 entries.
// Every entry is replicated per lane (word = tab[index * 32 + lane]), so a warp's 32 lookups hit 32 different
// banks whatever their indices: 8 KB per CTA, conflict-free.
constexpr int kSubstEntries = 64;
constexpr int kSubstWords = kSubstEntries * kWarp;

__device__ __forceinline__ void subst_table_fill(uint32_t* tab, const ShortConsts& K, int tid, int nthreads) {
    for (int e = tid; e < kSubstWords; e += nthreads) {
        const uint32_t idx = (uint32_t)e >> 5, ia = idx & 7u, ib = idx >> 3;
        const int lo = (ia & 3u) == 0u ? K.sm : K.sx, hi = (ib & 3u) == 0u ? K.sm : K.sx;
        tab[e] = ((uint32_t)hi << 16) | ((uint32_t)lo & 0xffffu);
    }
}
__device__ __forceinline__ uint32_t subst_row_part(uint32_t qa, uint32_t qb) { return (qa | (qb << 3)) << 7; }
__device__ __forceinline__ uint32_t subst_col_part(uint32_t ca, uint32_t cb) { return (((4u - ca) & 3u) | (((4u - cb) & 3u) << 3)) << 7; }
// row part (already holding the table's shared address and the lane's byte offset) + column part -> table word.
// The add is written as a multiply-add with an opaque 1 so that it issues on the fma pipe; the load is a plain
// (non-volatile) asm: the table never changes after the CTA's first barrier, the scheduler may hoist it freely.
__device__ __forceinline__ uint32_t subst_lookup(uint32_t row_part, uint32_t col_part, uint32_t one32) {
    uint32_t addr, v;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(row_part), "r"(one32), "r"(col_part));
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__host__ __device__ inline uint32_t dup16(int v) { return ((uint32_t)(uint16_t)(int16_t)v) * 0x10001u; }

__host__ inline ShortConsts make_short_consts(const Scores& sc, int type) {
    ShortConsts k;
    const int sm = 4 * (sc.match - sc.gap) + 1, sx = 4 * (sc.mismatch - sc.gap) + 1;
    k.tab_diff = ((uint32_t)(uint8_t)(int8_t)sm) ^ ((uint32_t)(uint8_t)(int8_t)sx);
    k.tab_mis = ((uint32_t)(uint8_t)(int8_t)sx) * 0x01010101u;
    k.cu = dup16(4 * sc.gap - 1);
    k.mask = 0xfffcfffcu; k.one = 0x00010001u; k.four = 4u;
    k.gap = sc.gap;
    k.init = (type == 0) ? sc.gap : 0;
    k.sm = sm; k.sx = sx; k.one32 = 1u;
    return k;
}

// PTX prmt in its default mode: selector nibble bit 3 replicates the sign of the chosen byte,
// which is how an int8 table entry becomes a sign-extended int16 half in one instruction.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t b, uint32_t c) {   // (a & b) | c
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ int half_lo(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
__device__ __forceinline__ int half_hi(uint32_t v) { return (int)(int16_t)(v >> 16); }
template <int HI> __device__ __forceinline__ int half_of(uint32_t v) { return HI ? half_hi(v) : half_lo(v); }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)hi << 16) | ((uint32_t)lo & 0xffffu); }

template <int N>
__device__ __forceinline__ uint32_t pick_any(const uint32_t (&Y)[N], uint32_t r) {   // Y[r] for a lane-varying r
    uint32_t v = Y[0];
#pragma unroll
    for (int k = 1; k < N; ++k) if (r == (uint32_t)k) v = Y[k];
    return v;
}

// packed maximum of the 32 registers (16 VIMNMX3.S16x2)
__device__ __forceinline__ uint32_t max_tree16(const uint32_t (&Y)[32]) {
    uint32_t m[12];
#pragma unroll
    for (int k = 0; k < 10; ++k) m[k] = __vimax3_s16x2(Y[3 * k], Y[3 * k + 1], Y[3 * k + 2]);
    m[10] = Y[30]; m[11] = Y[31];
    uint32_t n[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) n[k] = __vimax3_s16x2(m[3 * k], m[3 * k + 1], m[3 * k + 2]);
    return __vmaxs2(__vimax3_s16x2(n[0], n[1], n[2]), n[3]);
}
// the same for a shorter register block (N = 8, 16, 24): the trimmed last block of align_fill_short.cuh
template <int N>
__device__ __forceinline__ uint32_t max_tree16(const uint32_t (&Y)[N]) {
    static_assert(N == 8 || N == 16 || N == 24, "register block heights are multiples of 8");
    uint32_t m[N / 2];
#pragma unroll
    for (int k = 0; k < N / 4; ++k) {      // 4 registers -> 2 (one three-way, one passed on)
        m[2 * k] = __vimax3_s16x2(Y[4 * k], Y[4 * k + 1], Y[4 * k + 2]);
        m[2 * k + 1] = Y[4 * k + 3];
    }
    uint32_t v = m[0];
#pragma unroll
    for (int k = 1; k + 1 < N / 2; k += 2) v = __vimax3_s16x2(v, m[k], m[k + 1]);
    return __vmaxs2(v, m[N / 2 - 1]);
}

// For both halves at once: 31 - (the smallest r whose half of Y[r] equals that half of `cm`, the packed maximum of
// the 32 registers). Y[r] - cm is 0 where equal and at most -4 elsewhere (values carry the same tag, and a block's
// values lie within the 16-bit window of each other, so the packed subtraction cannot wrap): max(Y[r] - cm, -1) is a
// 0 / -1 mask, (mask & 0xffc0) | (31 - r) is 31 - r where equal and negative elsewhere, and a packed max tree picks
// the smallest r. 32 VIADDMNMX + 32 LOP3 + 16 VIMNMX3 for the two blocks of a lane.
__device__ __forceinline__ uint32_t first_rows_of_max(const uint32_t (&Y)[32], uint32_t cm) {
    const uint32_t negm = __vneg2(cm);
    const uint32_t keep = 0xffc0ffc0u;
    auto key = [&](int r) { return lop3_and_or(__viaddmax_s16x2(Y[r], negm, 0xffffffffu), keep, dup16(31 - r)); };
    uint32_t a = key(0), b = key(1);   // two running maxima (keys are folded as they are made: no 32-register array)
#pragma unroll
    for (int r = 2; r + 3 < 32; r += 4) { a = __vimax3_s16x2(a, key(r), key(r + 1)); b = __vimax3_s16x2(b, key(r + 2), key(r + 3)); }
    a = __vimax3_s16x2(a, key(30), key(31));
    return __vmaxs2(a, b);
}

// Direction word layout written by this kernel (read back by walk_kernel, klass kClassShort):
//   uint4 at dirs[dir_off + ((block * Tg + (j-1)) * 32 + lane) * 4 .. +3]; word k covers rows
//   8k..8k+7 of the block, low half = pair A, high half = pair B, row 8k in the top 2 bits of the
//   half. Stored value is the TAG (2 diagonal, 1 left, 0 up; 3 = local stop).
//
// Moving frame: the register value of cell (i,j) is Y = 4*H(i,j) - 4*gap*j + 1. In that frame a
// step to the right costs nothing, so the three candidates are
//     diagonal  Y(i-1,j-1) + 4*(s-gap) + 1   -> 4H' + 2
//     left      Y(i,  j-1)                    -> 4H' + 1
//     up        Y(i-1,j)   + 4*gap - 1        -> 4H' + 0
// (H' = H - gap*j) and one signed max gives maximum, tie order and direction tag at once.
// The state one thread carries through the row blocks of its two pairs; block<RB>() sweeps one block of RB
// (8, 16, 24 or 32) register rows over every column. Blocks are 32 rows apart whatever RB is: only the LAST
// block of a group is trimmed (150-row pairs: 4 x 32 + 24 rows instead of 5 x 32), its unused direction words
// are stored as zeros and never read (the walker only visits rows <= Q).
// SUB bit 0: substitution term by PRMT (0) or by shared-memory lookup (1, SubstTable).
// SUB bit 1: software-pipelined columns. The row loop of a column is one dependent chain (fused add-max, then the
//   re-arm LOP3: 8 cycles per row) with little else to issue next to it once the 32 independent "diagonal or left"
//   maxima of the column have been computed up front -- ncu shows the warps of a scheduler waiting on that chain
//   (stall "wait") while the alu pipe idles. Pipelined, the loop computes next column's maximum for row r right after
//   this column's value of row r exists: the independent work is spread along the chain instead of preceding it.
template <int TYPE, int SUBV>
struct ShortSweep {
    static constexpr int SUB = SUBV & 1;
    static constexpr bool PIPE = (SUBV & 2) != 0;
    uint32_t tab_at;       // SUB = 1: shared address of the lane's column of the substitution table
    // the thread's two pairs
    uint32_t QA, TA, QB, TB;
    const uint32_t *qwA, *twA, *qwB, *twB;
    bool liveA, liveB;
    uint32_t Tm, n_blocks;
    // per group
    uint32_t* dirs_g;      // dirs + dir_off + lane * 4, or nullptr
    uint32_t Tg;
    uint32_t* my_bnd;
    // results
    int resA, resB;
    int colbestA, colbestB, rowbestA, rowbestB;      // semiGlobal candidates (last column / row Q)
    uint32_t coliA, coliB, rowjA, rowjB;
    int bvA, bvB;                                    // local: running first maximum in row-major order
    uint32_t biA, bjA, biB, bjB;

    __device__ __forceinline__ void reset() {
        resA = 0; resB = 0;
        // last column: smallest i first, H(0,T) = 0 leads; last row: smallest j, H(Q,0) = 0 leads
        colbestA = 0; colbestB = 0; rowbestA = 0; rowbestB = 0;
        coliA = 0; coliB = 0; rowjA = 0; rowjB = 0;
        bvA = INT_MIN; bvB = INT_MIN;
        biA = 0; bjA = 0; biB = 0; bjB = 0;
    }

    template <int RB>
    __device__ __forceinline__ void block(const ShortConsts& K, const uint32_t b) {
        constexpr int R = kShortRows;   // block stride in rows
        const uint32_t MASK = K.mask, ONE = K.one, FOUR = K.four;
        const int frame = 4 * (K.init - K.gap);   // top border row in the moving frame: Y(0,j) = frame*j + 1
        const uint32_t i0 = b * R;   // rows i0+1 .. i0+RB
        // per-row PRMT selectors: byte0 = tabA[qA], byte1 = its sign, byte2 = tabB[qB], byte3 = sign
        uint32_t sel[RB], Y[RB];
        {
            // (sequence words through L2: they are fetched a step ahead anyway, and in streaming mode they were written
            // by a pack kernel that ran while this one was already resident -- L1 is not coherent with that)
            const uint32_t qa0 = __ldcg(qwA + (i0 >> 4)), qa1 = RB > 16 ? __ldcg(qwA + (i0 >> 4) + 1) : 0u;
            const uint32_t qb0 = __ldcg(qwB + (i0 >> 4)), qb1 = RB > 16 ? __ldcg(qwB + (i0 >> 4) + 1) : 0u;
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const uint32_t ca = ((r < 16 ? qa0 : qa1) >> (2 * (r & 15))) & 3u;
                const uint32_t cb = ((r < 16 ? qb0 : qb1) >> (2 * (r & 15))) & 3u;
                sel[r] = SUB ? tab_at + subst_row_part(ca, cb)
                             : ca | ((8u + ca) << 4) | ((4u + cb) << 8) | ((12u + cb) << 12);
                Y[r] = dup16(4 * (int)((i0 + 1 + r) * (uint32_t)K.init) + 1);   // column 0, frame 0
            }
        }
        uint32_t top_prev = dup16(4 * (int)(i0 * (uint32_t)K.init) + 1);       // Y(i0, 0)
        // software prefetch: the boundary row and the packed target words are fetched one step early
        // (two columns early: the compiler sinks the load to the end of the body, so a distance of one
        // column would leave only a few instructions between issue and use)
        uint32_t top_next = (b == 0) ? dup16(frame * 1 + 1) : my_bnd[(size_t)1 * kWarp];
        uint32_t top_next2 = (b == 0) ? dup16(frame * 2 + 1) : my_bnd[(size_t)2 * kWarp];
        uint32_t tA_next = __ldcg(twA), tB_next = __ldcg(twB), tA = 0, tB = 0;
        uint32_t* dcol = dirs_g ? dirs_g + (uint64_t)b * Tg * 128 : nullptr;

        // hoisted end-cell test: the column at which this block holds cell (Q,T) of either pair
        const uint32_t jhitA = (TYPE == 0 && liveA && (QA - 1) / R == b) ? TA : 0u;
        const uint32_t jhitB = (TYPE == 0 && liveB && (QB - 1) / R == b) ? TB : 0u;
        // rows of each pair inside this block, and whether the block holds the pair's row Q
        const uint32_t nvA = (liveA && QA > i0) ? min((uint32_t)RB, QA - i0) : 0u;
        const uint32_t nvB = (liveB && QB > i0) ? min((uint32_t)RB, QB - i0) : 0u;
        const bool lastA = nvA && QA <= i0 + R, lastB = nvB && QB <= i0 + R;
        const uint32_t rqA = lastA ? QA - 1 - i0 : 0u, rqB = lastB ? QB - 1 - i0 : 0u;
        const bool full_rows = nvA == (uint32_t)RB && nvB == (uint32_t)RB;
        const uint32_t Tmin = min(TA, TB);
        // codes and substitution tables of column jj (columns are visited in order: a 16-column word at a time)
        uint32_t tabA = 0, tabB = 0;
        auto column_tables = [&](uint32_t jj) {
            if (((jj - 1) & 15u) == 0) {
                tA = tA_next; tB = tB_next;
                tA_next = __ldcg(twA + ((jj - 1) >> 4) + 1); tB_next = __ldcg(twB + ((jj - 1) >> 4) + 1);
            }
            const uint32_t cA = tA & 3u, cB = tB & 3u;
            tA >>= 2; tB >>= 2;
            // per-column byte tables: entry c = S'(c, target) (match where c == target code); or the column part
            // of the shared-memory table index
            tabA = SUB ? subst_col_part(cA, cB) : K.tab_mis ^ (K.tab_diff << (8 * cA));
            tabB = SUB ? K.one32 : K.tab_mis ^ (K.tab_diff << (8 * cB));
        };
        auto subst = [&](int r) { return SUB ? subst_lookup(sel[r], tabA, tabB) : prmt(tabA, tabB, sel[r]); };
        uint32_t M[PIPE ? RB : 1];   // PIPE: max(diagonal + S, left) of the column about to be swept
        if (PIPE) {
            column_tables(1);
#pragma unroll
            for (int r = 0; r < RB; ++r) M[r] = __viaddmax_s16x2(r ? Y[r - 1] : top_prev, subst(r), Y[r]);
        }
#pragma unroll 2
        for (uint32_t j = 1; j <= Tm; ++j) {
            // PIPE: the tables of the NEXT column (its maxima are made in this sweep; one word past the end is a spare)
            column_tables(PIPE ? j + 1 : j);
            const uint32_t top = top_next;
            top_next = top_next2;
            if (j + 2 <= Tm) top_next2 = (b == 0) ? dup16(frame * (int)(j + 2) + 1) : my_bnd[(size_t)(j + 2) * kWarp];
            uint32_t up = top;        // Y of the row above, this column's frame
            uint32_t dg = top_prev;   // Y(i0, j-1), previous column's frame
            top_prev = top;
            uint32_t accZ = 0, accY = 0, w[4] = {0u, 0u, 0u, 0u};
            const uint32_t clampv = dup16(3 - 4 * K.gap * (int)j);   // local: H = 0 with the stop tag, this column's frame
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const uint32_t above = up;
                const uint32_t m1 = PIPE ? M[r] : __viaddmax_s16x2(dg, subst(r), Y[r]);
                uint32_t Z = __viaddmax_s16x2(above, K.cu, m1);
                if (TYPE == 1) Z = __vmaxs2(Z, clampv);   // clamp at 0 (team_alignment.cpp:185), tag 3 = stop
                dg = Y[r];
                Y[r] = lop3_and_or(Z, MASK, ONE);
                up = Y[r];
                if (PIPE) M[r] = __viaddmax_s16x2(above, subst(r), Y[r]);   // next column: diagonal = the row above, left = this cell
                // direction tags: two multiply-add chains on the fma pipe; (Z - Y) = tag - 1 per half
                accZ = accZ * FOUR + Z;
                accY = accY * FOUR + Y[r];
                if ((r & 7) == 7) { w[r >> 3] = accZ - accY + 0x55555555u; accZ = 0; accY = 0; }
            }
            if (b + 1 < n_blocks) my_bnd[(size_t)j * kWarp] = up;   // (RB = 32 whenever a block follows)
            // streaming store: the direction matrix is written once and must not evict the boundary rows from L2
            if (dcol) __stcs(reinterpret_cast<uint4*>(dcol + (uint64_t)(j - 1) * 128), make_uint4(w[0], w[1], w[2], w[3]));
            // end-cell capture (global): the cell (Q, T) of either pair
            if (TYPE == 0) {
                const bool hitA = j == jhitA, hitB = j == jhitB;
                if (hitA || hitB) {
                    const uint32_t rA = (QA - 1) % R, rB = (QB - 1) % R;
                    const uint32_t vA = pick_any(Y, rA), vB = pick_any(Y, rB);
                    if (hitA) resA = (half_lo(vA) - 1 + 4 * K.gap * (int)j) >> 2;
                    if (hitB) resB = (half_hi(vB) - 1 + 4 * K.gap * (int)j) >> 2;
                }
            }
            const int back = 4 * K.gap * (int)j - 1;   // H = (Y + back) >> 2 in this column's frame
            if (TYPE == 2) {
                if ((j == TA && nvA) || (j == TB && nvB)) {   // a pair's last column: rows ascend, strict '>' keeps the smallest i
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        const int hA = (half_lo(Y[r]) + back) >> 2, hB = (half_hi(Y[r]) + back) >> 2;
                        if (j == TA && (uint32_t)r < nvA && hA > colbestA) { colbestA = hA; coliA = i0 + 1 + r; }
                        if (j == TB && (uint32_t)r < nvB && hB > colbestB) { colbestB = hB; coliB = i0 + 1 + r; }
                    }
                }
                if (lastA || lastB) {   // row Q of a pair, every column: strict '>' keeps the smallest j
                    const int hA = (half_lo(pick_any(Y, rqA)) + back) >> 2, hB = (half_hi(pick_any(Y, rqB)) + back) >> 2;
                    if (lastA && j <= TA && hA > rowbestA) { rowbestA = hA; rowjA = j; }
                    if (lastB && j <= TB && hB > rowbestB) { rowbestB = hB; rowjB = j; }
                }
            }
            if (TYPE == 1) {
                int mA, mB;   // column maxima over the valid rows, as register halves
                uint32_t cm = 0;
                const bool packed_cm = full_rows && j <= Tmin;
                if (packed_cm) {
                    cm = max_tree16(Y);
                    mA = half_lo(cm); mB = half_hi(cm);
                } else {
                    mA = INT_MIN; mB = INT_MIN;
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        if ((uint32_t)r < nvA) mA = max(mA, half_lo(Y[r]));
                        if ((uint32_t)r < nvB) mB = max(mB, half_hi(Y[r]));
                    }
                }
                const int hA = (nvA && j <= TA) ? (mA + back) >> 2 : INT_MIN;
                const int hB = (nvB && j <= TB) ? (mB + back) >> 2 : INT_MIN;
                // a new maximum, or a tie that may sit on a smaller row of this block than the current holder
                const bool trigA = hA > bvA || (hA == bvA && biA > i0 + 1);
                const bool trigB = hB > bvB || (hB == bvB && biB > i0 + 1);
                if (trigA || trigB) {
                    uint32_t rA = RB, rB = RB;   // smallest row of the block that holds each pair's column maximum
                    bool found = false;
                    if constexpr (RB == 32) {
                        // full blocks: both pairs' rows at once with packed arithmetic (the scalar search below costs a
                        // compare and a select per row and pair, and on related pairs it runs on every diagonal column)
                        if (packed_cm && (cm & 0xffffu) != 0x8000u && (cm >> 16) != 0x8000u) {   // (-32768 has no packed negative)
                            const uint32_t km = first_rows_of_max(Y, cm);
                            rA = 31u - (uint32_t)half_lo(km); rB = 31u - (uint32_t)half_hi(km);
                            found = true;
                        }
                    }
                    if (!found) {
#pragma unroll
                        for (int r = RB - 1; r >= 0; --r) {
                            if (trigA && (uint32_t)r < nvA && half_lo(Y[r]) == mA) rA = r;
                            if (trigB && (uint32_t)r < nvB && half_hi(Y[r]) == mB) rB = r;
                        }
                    }
                    if (trigA && (hA > bvA || i0 + 1 + rA < biA)) { bvA = hA; biA = i0 + 1 + rA; bjA = j; }
                    if (trigB && (hB > bvB || i0 + 1 + rB < biB)) { bvB = hB; biB = i0 + 1 + rB; bjB = j; }
                }
            }
        }
    }
};

// Streaming mode (host pipeline of a uniform batch): ONE launch of the fill serves the whole batch while its bytes
// are still arriving. The upload goes in slices; after a slice has landed and been 2-bit packed, a one-thread kernel
// raises `watermark` to the number of pairs that are ready. A warp that has taken group g waits until its 64 pairs are
// below the watermark, sweeps them, and counts the group as done for its wave (= upload slice), on which that wave's
// traceback kernel waits. Warps of a persistent launch drift apart and keep the alu pipe busy; a launch per wave
// starts and ends every warp together and lost a third of the throughput. watermark == nullptr: the classic mode.
struct ShortStream {
    const uint32_t* watermark;   // pairs uploaded + packed so far (release / acquire)
    uint32_t* wave_done;         // [n_waves] groups finished, one counter per wave
    uint32_t* stall_flag;        // raised when a warp gives up waiting (never hang the device)
    uint32_t n_waves;
    uint32_t group_start[33];    // first group of each wave, then the group count
    uint64_t dir_base[32];       // word offset of each wave's direction block
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The trimmed variants are real calls: inlined next to the 32-row loop they cost that loop its register
// allocation (measured: -8 % on every block); a call per group is free.
template <int TYPE, int SUB, int RB>
__device__ __noinline__ void short_block_trimmed(ShortSweep<TYPE, SUB>* sw, const ShortConsts* K, uint32_t b) {
    ShortSweep<TYPE, SUB> s = *sw;   // by value: nothing the loop reads may alias its stores
    const ShortConsts k = *K;
    s.template block<RB>(k, b);
    *sw = s;
}

template <int TYPE, int SUB>
__global__ void __launch_bounds__(kShortThreads, 7)
fill_short_kernel(const uint32_t* __restrict__ qpk, const uint32_t* __restrict__ tpk,
                  const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ work, uint32_t n_work,
                  const ShortGroup* __restrict__ groups, uint32_t* __restrict__ group_counter,
                  const uint8_t* flags, ShortConsts K,
                  uint32_t* __restrict__ dirs, uint32_t* bnd, uint32_t bnd_cols,
                  int32_t* __restrict__ score, uint32_t* __restrict__ end_i, uint32_t* __restrict__ end_j,
                  const ShortStream stream_ctl) {
    constexpr int R = kShortRows;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_groups = (n_work + 63) / 64;
    ShortSweep<TYPE, SUB> sw;
    sw.my_bnd = bnd + (size_t)warp_global * bnd_cols * kWarp + lane;   // [col][lane]
    sw.tab_at = 0;
    if (SUB & 1) {
        __shared__ uint32_t subst_tab[kSubstWords];
        subst_table_fill(subst_tab, K, threadIdx.x, kShortThreads);
        __syncthreads();
        sw.tab_at = (uint32_t)__cvta_generic_to_shared(subst_tab) + (uint32_t)lane * 4u;
    }

    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(group_counter, 1u);
        g = __shfl_sync(kFull, g, 0);
        if (g >= n_groups) break;
        uint32_t wave = 0;
        uint64_t dir_base = 0;
        if (stream_ctl.watermark) {   // streaming mode: wait until the group's pairs have landed and been packed
            while (wave + 1 < stream_ctl.n_waves && g >= stream_ctl.group_start[wave + 1]) ++wave;
            dir_base = stream_ctl.dir_base[wave];
            const uint32_t need = min(n_work, (g + 1) * 64);
            uint32_t ok = 1;
            if (lane == 0) {
                uint32_t spins = 0;
                while (ld_acquire_gpu(stream_ctl.watermark) < need) {
                    __nanosleep(200);
                    if (++spins > (1u << 23)) { atomicExch(stream_ctl.stall_flag, 1u); ok = 0; break; }
                }
            }
            ok = __shfl_sync(kFull, ok, 0);
            if (!ok) break;
        }

        // my two pairs
        const uint32_t wa = g * 64 + 2 * lane, wb = wa + 1;
        uint32_t pA = 0xffffffffu, pB = 0xffffffffu;
        sw.QA = 0; sw.TA = 0; sw.QB = 0; sw.TB = 0;
        sw.qwA = qpk; sw.twA = tpk; sw.qwB = qpk; sw.twB = tpk;
        sw.Tg = groups[g].cols;
        sw.dirs_g = dirs ? dirs + dir_base + groups[g].dir_off + (uint64_t)lane * 4 : nullptr;
        if (wa < n_work) {
            pA = work[wa];
            const PairDesc d = pairs[pA];
            if (__ldcg(flags + pA) == 0) { sw.QA = d.Q; sw.TA = d.T; sw.qwA = qpk + d.qpk_off; sw.twA = tpk + d.tpk_off; }
            else pA = 0xffffffffu;   // not pure ACGT: the generic kernel owns it
        }
        if (wb < n_work) {
            pB = work[wb];
            const PairDesc d = pairs[pB];
            if (__ldcg(flags + pB) == 0) { sw.QB = d.Q; sw.TB = d.T; sw.qwB = qpk + d.qpk_off; sw.twB = tpk + d.tpk_off; }
            else pB = 0xffffffffu;
        }
        const uint32_t QA = sw.QA, TA = sw.TA, QB = sw.QB, TB = sw.TB;
        const bool liveA = QA && TA, liveB = QB && TB;   // has inner cells
        sw.liveA = liveA; sw.liveB = liveB;
        const uint32_t Qm = max(liveA ? QA : 0u, liveB ? QB : 0u);
        sw.Tm = max(liveA ? TA : 0u, liveB ? TB : 0u);
        sw.n_blocks = (Qm + R - 1) / R;
        sw.reset();
        // The height of a block is a warp-wide decision (lanes must not diverge over the variants): full blocks
        // everywhere except the last block of the group's longest query, which is trimmed to a multiple of 8 rows.
        const uint32_t Qg = __reduce_max_sync(kFull, Qm);
        const uint32_t nbg = (Qg + R - 1) / R;
        const uint32_t last_rows = (Qg - (nbg ? nbg - 1 : 0u) * R + 7u) & ~7u;   // 8 .. 32 (0 when the group is empty)

        for (uint32_t b = 0; b < sw.n_blocks; ++b) {
            if (b + 1 < nbg || last_rows == 32u) sw.template block<32>(K, b);
            else if (last_rows == 24u) short_block_trimmed<TYPE, SUB, 24>(&sw, &K, b);
            else if (last_rows == 16u) short_block_trimmed<TYPE, SUB, 16>(&sw, &K, b);
            else short_block_trimmed<TYPE, SUB, 8>(&sw, &K, b);
        }
        if (TYPE == 0) {
            if (pA != 0xffffffffu) {
                score[pA] = liveA ? sw.resA : (int)((QA + TA) * (uint32_t)K.init);
                end_i[pA] = QA; end_j[pA] = TA;
            }
            if (pB != 0xffffffffu) {
                score[pB] = liveB ? sw.resB : (int)((QB + TB) * (uint32_t)K.init);
                end_i[pB] = QB; end_j[pB] = TB;
            }
        }
        if (TYPE == 2) {   // last column wins ties, the last row only if strictly greater (team_alignment.cpp:265-278)
            if (pA != 0xffffffffu) {
                if (!liveA) { score[pA] = 0; end_i[pA] = 0; end_j[pA] = TA; }
                else if (sw.rowbestA > sw.colbestA) { score[pA] = sw.rowbestA; end_i[pA] = QA; end_j[pA] = sw.rowjA; }
                else { score[pA] = sw.colbestA; end_i[pA] = sw.coliA; end_j[pA] = TA; }
            }
            if (pB != 0xffffffffu) {
                if (!liveB) { score[pB] = 0; end_i[pB] = 0; end_j[pB] = TB; }
                else if (sw.rowbestB > sw.colbestB) { score[pB] = sw.rowbestB; end_i[pB] = QB; end_j[pB] = sw.rowjB; }
                else { score[pB] = sw.colbestB; end_i[pB] = sw.coliB; end_j[pB] = TB; }
            }
        }
        if (TYPE == 1) {
            if (pA != 0xffffffffu) { score[pA] = liveA ? sw.bvA : 0; end_i[pA] = liveA ? sw.biA : 0u; end_j[pA] = liveA ? sw.bjA : 0u; }
            if (pB != 0xffffffffu) { score[pB] = liveB ? sw.bvB : 0; end_i[pB] = liveB ? sw.biB : 0u; end_j[pB] = liveB ? sw.bjB : 0u; }
        }
        if (stream_ctl.watermark) {
            // every lane's direction words and results device-wide before the wave's count (the traceback kernel of the
            // wave acquires the count): each lane fences its own stores, the warp barrier orders them before lane 0's
            __threadfence();
            __syncwarp();
            if (lane == 0) { __threadfence(); atomicAdd(stream_ctl.wave_done + wave, 1u); }
        }
    }
}

}  // namespace b200
