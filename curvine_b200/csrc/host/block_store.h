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
};
Err parse_data_dir(const std::string& spec, StorageDir* out);  // worker_conf.rs:59-95

struct BlockMeta {
    int64_t id = 0;
    int64_t len = 0;
    int32_t storage_type = kStorageDisk;
    std::string path;
};

class BlockStore {
   public:
    Err init(const std::vector<std::string>& data_dirs, const std::string& cluster_id);
    Err get_block(int64_t id, BlockMeta* out) const;  // FsError::Common "block N not exits" style on miss
    // create/overwrite a finalized block file in the next dir of `storage_type` (round robin, policy.rs:56-105)
    Err put_block(int64_t id, const void* data, int64_t len, int32_t storage_type, std::string* path_out);
    Err register_block(int64_t id, int64_t len, int32_t storage_type, const std::string& path);
    // writer side: directory + path a new block of `storage_type` is written to (BlockStore::open_block)
    Err open_block_path(int64_t id, int32_t storage_type, std::string* path_out, int32_t* dir_storage_type);
    void remove_block(int64_t id);
    size_t num_blocks() const;
    const std::vector<StorageDir>& dirs() const { return dirs_; }
    // pick the directory a new block of `storage_type` goes to (falls back to Disk dirs, then any)
    const StorageDir* choose_dir(int32_t storage_type);

   private:
    Err scan_dir(const StorageDir& d);
    std::vector<StorageDir> dirs_;
    mutable std::mutex mu_;
    std::unordered_map<int64_t, BlockMeta> blocks_;
    size_t rr_ = 0;
};

}  // namespace cv
