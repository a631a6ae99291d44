// Worker BlockStore: directory layout, block-id packing, block lookup.
//
// Mirrors (reference):
//   curvine-server/src/master/meta/inode_id.rs:22-60          block_id = (inode & (2^40-1)) << 24 | seq
//   curvine-server/src/worker/block/block_meta.rs:44-46,199-237  <base>/active/b{(id>>48)&31}/b{(id>>32)&31}/blk_<id>
//   curvine-server/src/worker/storage/vfs_dataset.rs:132-157, vfs_dir.rs:339-362  rescan active/ on start
//   curvine-common/src/conf/worker_conf.rs:59-95             "[MEM:10MB]/path" dir tags
//   curvine-server/src/worker/block/block_store.rs:59-63      get_block
// Block files are raw bytes (no header/footer/checksum).
#pragma once
#include <mutex>
#include <unordered_map>
#include <vector>

#include <memory>

#include "arena.h"
#include "common.h"
#include "wire.h"

namespace cv {

constexpr int64_t kInodeIdMask = (1ll << 40) - 1;
constexpr int64_t kSeqMask = (1ll << 24) - 1;

Err create_block_id(int64_t inode_id, int64_t seq, int64_t* out);
inline int64_t block_inode(int64_t id) { return (id >> 24) & kInodeIdMask; }
inline int64_t block_seq(int64_t id) { return id & kSeqMask; }
std::string block_dir(const std::string& base, int64_t id);
std::string block_path(const std::string& base, int64_t id);

struct StorageDir {
    int32_t storage_type = kStorageDisk;
    int64_t capacity = 0;
    std::string path;       // as configured
    std::string base_path;  // path/<cluster_id>
    std::shared_ptr<MemArena> arena;  // [worker] mem_arena: this MEM dir keeps its blocks as extents of one pinned-once arena
};

// [worker] mem_arena / arena_segment / arena_numa
struct ArenaOpts {
    bool enable = false;
    int64_t seg_bytes = 1ll << 30;
    int64_t reuse_delay_ms = 1000;  // quarantine of freed extents (arena.h)
    std::vector<int> numa;  // NUMA node per MEM data dir, in data_dir order (-1 / missing: wherever the worker runs)
};
Err parse_data_dir(const std::string& spec, StorageDir* out);  // worker_conf.rs:59-95

struct BlockMeta {
    int64_t id = 0;
    int64_t len = 0;
    int32_t storage_type = kStorageDisk;
    std::string path;  // the block file; for an arena block: the arena segment file that holds the extent
    std::shared_ptr<ExtentHold> hold;  // arena block: lives at [hold->ext.off, +len) of `path`; copies of the meta keep it allocated
    bool in_arena() const { return hold != nullptr; }
    const uint8_t* mem() const { return hold ? hold->ptr() : nullptr; }  // worker-side mapping of the bytes
};

// Where a block being written goes (BlockStore::open_block): a file to pwrite, or an arena extent to copy into.
struct BlockWriteTarget {
    std::string path;
    int32_t dir_storage_type = kStorageDisk;
    std::shared_ptr<MemArena> arena;
    ArenaExtent ext;
    std::string stub_path;  // arena: the reference-layout path that will hold the extent descriptor
    uint8_t* mem() const { return arena ? arena->ptr(ext) : nullptr; }
};

class BlockStore {
   public:
    Err init(const std::vector<std::string>& data_dirs, const std::string& cluster_id, const ArenaOpts& arena = ArenaOpts());
    Err get_block(int64_t id, BlockMeta* out) const;  // FsError::Common "block N not exits" style on miss
    // create/overwrite a finalized block file in the next dir of `storage_type` (round robin, policy.rs:56-105)
    // dir_hint >= 0 pins the choice to the (dir_hint mod n)-th dir of that type instead of the shared round-robin cursor
    // (deterministic placement: block b of a striped file -> the arena next to the GPU that will ingest it)
    Err put_block(int64_t id, const void* data, int64_t len, int32_t storage_type, std::string* path_out, int dir_hint = -1);
    // same, but the caller fills the bytes in place (arena dirs only hand out memory; file dirs return mem == nullptr)
    Err reserve_block(int64_t id, int64_t len, int32_t storage_type, int dir_hint, BlockWriteTarget* out);
    // writer side (BlockStore::open_block): where a block of `storage_type`, at most `block_size` long, is written.  Re-opening
    // a finalized block writes to the same file; an arena block gets a fresh extent holding a copy of its bytes.
    Err open_block(int64_t id, int32_t storage_type, int64_t block_size, BlockWriteTarget* out);
    Err commit_block(int64_t id, BlockWriteTarget* t, int64_t len);  // finalize at `len` bytes
    void abort_block(int64_t id, BlockWriteTarget* t);
    void remove_block(int64_t id);  // drops the block: unlinks its file / frees its extent and descriptor
    size_t num_blocks() const;
    const std::vector<StorageDir>& dirs() const { return dirs_; }
    // pick the directory a new block of `storage_type` goes to (falls back to Disk dirs, then any)
    const StorageDir* choose_dir(int32_t storage_type, int dir_hint = -1);

   private:
    Err register_meta(const BlockMeta& m);
    Err scan_dir(const StorageDir& d);
    std::vector<StorageDir> dirs_;
    mutable std::mutex mu_;
    std::unordered_map<int64_t, BlockMeta> blocks_;
    // arena blocks that are open for writing and not committed yet: the worker builds a fresh handler for every non-Running
    // message (worker_handler.rs:71-88), so the Complete of a write finds its extent here -- the counterpart of the reference
    // re-deriving the block file path from the id
    std::unordered_map<int64_t, BlockWriteTarget> writing_;
    size_t rr_ = 0;
};

}  // namespace cv
