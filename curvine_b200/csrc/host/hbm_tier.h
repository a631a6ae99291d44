// HBM as a worker tier (the worker-side CUDA counterpart of ReadHandler): blocks resident in device
// memory are served to remote clients as orpc frames that were PACKED ON THE GPU (K4: prefixes written, payload copied
// behind them, CRC computed at the source) and moved D2H once per block.
//
// Reference shape it mirrors: ReadHandler::{open,read} (curvine-server/src/worker/handler/read_handler.rs:60-183) +
// RpcFrame::send (orpc/src/handler/rpc_frame.rs:205-220); the tier sits beside Mem/Ssd/Hdd
// (curvine-common/src/state/storage_info.rs:36-49).
#pragma once
#include <memory>
#include <mutex>
#include <unordered_map>

#include "common.h"

namespace cv {

struct HbmBlock {
    uint8_t* d_ptr = nullptr;
    int64_t len = 0;
    int device = 0;
};

// The packed response stream of one block read: frame f = wire + f*(22+chunk), carries min(chunk, remaining) bytes.
struct PackedStream {
    uint8_t* wire = nullptr;  // pinned host memory
    size_t wire_cap = 0;
    int64_t off0 = 0, total = 0, chunk = 0;
    int64_t req_id = 0;
    int32_t first_seq = 1;
    uint32_t crc32c = 0;  // CRC-32C of the packed payload, computed at the source by K4
    ~PackedStream();
};

class HbmTier {
   public:
    ~HbmTier();
    Err load(int64_t block_id, const void* host_bytes, int64_t len, int device);
    bool get(int64_t block_id, HbmBlock* out) const;
    size_t size() const;
    // K4 over [off, off+n) of a resident block: prefixes (code 81, status Running|Success, req_id, seq first_seq..) + payload
    Err pack(const HbmBlock& b, int64_t off, int64_t n, int64_t chunk, int64_t req_id, int32_t first_seq, PackedStream* out) const;

   private:
    mutable std::mutex mu_;
    std::unordered_map<int64_t, HbmBlock> blocks_;
};

}  // namespace cv
