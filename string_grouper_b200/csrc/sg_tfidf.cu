// K1 placeholder — replaced by the real kernels in the next commit.
#include "sg_common.cuh"
using namespace sg;
extern "C" {
int64_t sg_tfidf_table_slots(int ngram) { return ngram >= 1 && ngram <= 4 ? (int64_t)1 << (7 * ngram) : -1; }
int sg_tfidf_count(const uint8_t *, const int64_t *, int64_t, int, unsigned, int32_t *, uint32_t *, uint16_t *, int32_t *, void *) { return fail(SG_ERR_UNSUPPORTED, "K1 not built yet"); }
size_t sg_tfidf_finalize_workspace_bytes(int64_t, int) { return 0; }
int sg_tfidf_finalize(const int64_t *, int64_t, int, int, int32_t *, const uint32_t *, const uint16_t *, const int32_t *, int64_t *, int32_t *, double *, float *, int32_t *, int64_t *, void *, size_t, void *) { return fail(SG_ERR_UNSUPPORTED, "K1 not built yet"); }
int sg_tfidf_vocab_keys(const int32_t *, int, const int32_t *, uint32_t *, void *) { return fail(SG_ERR_UNSUPPORTED, "K1 not built yet"); }
}
