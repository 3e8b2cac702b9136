// decode_host.inl -- placeholder until the decode kernels land (next commit).
extern "C"
{
    int32_t qoipp_b200_decode_dev(qoipp_b200_ctx*, const uint8_t*, uint64_t, const qoipp_b200_desc*, uint8_t, int32_t, uint8_t*, uint64_t, void*) { return -(int32_t)cudaErrorNotSupported; }
    int32_t qoipp_b200_decode_status(qoipp_b200_ctx*, void*, int32_t*) { return -(int32_t)cudaErrorNotSupported; }
    int32_t qoipp_b200_decode_host(qoipp_b200_ctx*, const uint8_t*, uint64_t, uint8_t, int32_t, uint8_t*, uint64_t, qoipp_b200_desc*) { return -(int32_t)cudaErrorNotSupported; }
    int32_t qoipp_b200_decode_batch_dev(qoipp_b200_ctx*, const uint8_t*, const uint64_t*, uint32_t, const qoipp_b200_desc*, uint8_t, uint8_t*, uint64_t, void*) { return -(int32_t)cudaErrorNotSupported; }
    int32_t qoipp_b200_stream_decode_host(qoipp_b200_ctx*, qoipp_b200_state*, const uint8_t*, uint64_t, uint8_t*, uint64_t, uint64_t*, uint64_t*) { return -(int32_t)cudaErrorNotSupported; }
}
