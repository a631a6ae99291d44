"""Worker BlockStore on-disk layout + block-id packing.

Oracle / test infrastructure only (see oracle/__init__.py).

Follows:
  * curvine-server/src/master/meta/inode_id.rs:22-60  block_id = (inode & (2^40-1)) << 24 | seq
  * curvine-server/src/worker/block/block_meta.rs:44-46,199-237
        <base>/active/b{(id>>48)&0x1F}/b{(id>>32)&0x1F}/blk_<id>   (Finalized / Writing)
        <base>/staging/blk_<id>                                     (Recovering)
  * curvine-server/src/worker/storage/mod.rs:42-45   ACTIVE_DIR="active", STAGING_DIR="staging"
  * block files are raw bytes: no header, footer or checksum.
``base`` = <data_dir>/<cluster_id> (vfs_dir.rs).
"""
import os

ID_BITS = 40
SEQ_BITS = 24
ID_MASK = (1 << ID_BITS) - 1
SEQ_MASK = (1 << SEQ_BITS) - 1
ACTIVE_DIR = "active"
STAGING_DIR = "staging"


def create_block_id(inode_id: int, seq: int) -> int:
    if inode_id > ID_MASK:
        raise ValueError("inode id exceeds maximum value %d" % ID_MASK)
    if seq > SEQ_MASK:
        raise ValueError("seq id exceeds maximum value %d" % SEQ_MASK)
    return ((inode_id & ID_MASK) << SEQ_BITS) | (seq & SEQ_MASK)


def block_inode(block_id: int) -> int:
    return (block_id >> SEQ_BITS) & ID_MASK


def block_seq(block_id: int) -> int:
    return block_id & SEQ_MASK


def block_dir(base: str, block_id: int, recovering: bool = False) -> str:
    if recovering:
        return os.path.join(base, STAGING_DIR)
    uid = block_id & 0xFFFFFFFFFFFFFFFF
    return os.path.join(base, ACTIVE_DIR, "b%d" % ((uid >> 48) & 0x1F), "b%d" % ((uid >> 32) & 0x1F))


def block_path(base: str, block_id: int, recovering: bool = False) -> str:
    return os.path.join(block_dir(base, block_id, recovering), "blk_%d" % block_id)
