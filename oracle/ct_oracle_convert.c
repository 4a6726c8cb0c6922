/*
 * ct_oracle_convert.c -- TEST INFRASTRUCTURE ONLY (part of libct_oracle.so, see ct_oracle.c header).
 *
 * CPU restatement of the AutoAWQ -> compressed-tensors repack (reference:
 * /root/reference/src/compressed_tensors/entrypoints/convert/converters/autoawq.py:109-129, 179-262), written the long way the
 * reference does it -- unpack every nibble, undo AutoAWQ's order, mask, subtract 8, transpose, pack_to_int32 -- so that the
 * one-pass kernel is checked against the real chain and not against a copy of its own shortcut.
 * Pinned by tests/golden/convert.pt.gz (tests/golden/make_golden_convert.py).
 */

static const int ORC_AWQ_REVERSE[8] = {0, 4, 1, 5, 2, 6, 3, 7};

/* qweight int32 [K, N/8] -> int8 codes [N, K]: unpack_awq (:219-241), reverse_awq_order (:243-262), & 15, - 8, .T (:196-214) */
static void awq_codes(const int32_t* q, int8_t* codes, int64_t K, int64_t N) {
    int64_t NW = N / 8;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < K; ++k)
        for (int64_t w = 0; w < NW; ++w) {
            uint32_t word = (uint32_t)q[k * NW + w];
            int8_t un[8];
            for (int j = 0; j < 8; ++j) un[j] = (int8_t)(int32_t)(word >> (4 * j));       /* shift, .to(int8) keeps the low byte */
            for (int c = 0; c < 8; ++c) {
                int8_t v = (int8_t)((un[ORC_AWQ_REVERSE[c]] & 15) - 8);
                codes[(8 * w + c) * K + k] = v;                                            /* transposed: [N, K] */
            }
        }
}

int orc_awq_repack(const int32_t* qweight, int32_t* out, int64_t K, int64_t N) {
    if (N % 8) return ORC_E_SHAPE;
    int8_t* codes = (int8_t*)malloc((size_t)(K * N ? K * N : 1));
    if (!codes) return ORC_E_SHAPE;
    awq_codes(qweight, codes, K, N);
    int rc = orc_pack_int32(codes, out, N, K, 4, 1);     /* pack_to_int32(weight, 4): [N, ceil(K/8)] */
    free(codes);
    return rc;
}

/* qzeros int32 [G, N/8] -> pack_to_int32(zero_point [N, G], 4, packed_dim=0).contiguous(): int32 [N/8, G]  (:124-128) */
int orc_awq_repack_zeros(const int32_t* qzeros, int32_t* out, int64_t G, int64_t N) {
    if (N % 8) return ORC_E_SHAPE;
    int8_t* codes = (int8_t*)malloc((size_t)(G * N ? G * N : 1));
    if (!codes) return ORC_E_SHAPE;
    awq_codes(qzeros, codes, G, N);                       /* [N, G] */
    int rc = orc_pack_int32(codes, out, N, G, 4, 0);      /* packs down dim 0 -> [N/8, G] row-major */
    free(codes);
    return rc;
}
