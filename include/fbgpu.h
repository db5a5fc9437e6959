/*
 * fbgpu.h — C ABI of libfbgpu: a B200-native (sm_100a) roaring-bitmap query executor that replaces the
 * per-shard inner loop of FeatureBase's executor.go / fragment.go / roaring/.
 *
 * The reference has no FFI seam on this path (it is 100 % Go).  Each entry point below names the
 * reference function(s) whose per-shard map step + reduce it replaces; a Go maintainer binds them through
 * cgo from executor.mapperLocal (INTEGRATION.md shows the stub).  Paths are relative to /root/reference.
 *
 * Conventions: all integers fixed width, little endian; no callbacks; inputs are read-only and never
 * retained after return (cgo pointer rules); outputs are caller-allocated; return 0 on success or a
 * negative FBGPU_E_* (maps to Go `error`; every executor function returns (T, error)).  Thread-safe and
 * re-entrant: any OS thread may call any function on a context (goroutines migrate between threads), the
 * library sets the CUDA device itself on every call.  One context per GPU; a context owns the contiguous
 * shard range its caller loads into it.
 */
#ifndef FBGPU_H
#define FBGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FBGPU_ABI_VERSION 2

/* error codes */
#define FBGPU_OK 0
#define FBGPU_E_INVALID (-1)  /* bad argument / malformed program                       */
#define FBGPU_E_QUERY (-2)    /* query the reference rejects (e.g. empty Intersect())    */
#define FBGPU_E_FORMAT (-3)   /* unreadable roaring data                                 */
#define FBGPU_E_NOSPACE (-4)  /* output buffer too small; required size reported         */
#define FBGPU_E_CUDA (-5)     /* CUDA runtime failure (see fbgpu_last_error)             */
#define FBGPU_E_NOMEM (-6)
#define FBGPU_E_COMM (-7)     /* NCCL failure / communicator not initialised             */

typedef struct fbgpu_ctx fbgpu_ctx; /* opaque, one per GPU */

/* fbgpu_init(FBGPU_DEVICE_NONE, ..) creates an INSPECTION-ONLY context: it touches no device, accepts the residency calls
 * (load / drop / commit / stats) and fbgpu_debug_container(), and refuses every query with FBGPU_E_CUDA.  It exists so that the
 * loaders and the store tables can be tested on a machine without a GPU; it is not a CPU execution path. */
#define FBGPU_DEVICE_NONE (-1)

/* lifecycle (what Holder.Open / Holder.Close are to the fragments of a node, holder.go:432, 614): one context per GPU,
 * created once per process.  device_ordinal is the CUDA ordinal this context owns. */
int fbgpu_init(int32_t device_ordinal, fbgpu_ctx **out);
void fbgpu_shutdown(fbgpu_ctx *ctx);
/* thread-local message of the last failing call on this thread */
const char *fbgpu_last_error(void);
int32_t fbgpu_abi_version(void);

/* ---- residency (replaces fragment.row -> tx.OffsetRange -> rbf cursor walk: fragment.go:283-333,
 *      rbf.go:472, rbf/tx.go:1586-1637; data source = tx.RoaringBitmap / Bitmap.WriteTo bytes, tx.go:85) ----
 * `roaring` = Pilosa-roaring bytes (cookie 12348, roaring/roaring.go:1738-1817) or official RoaringBitmap
 * bytes (12346/12347, roaring.go:6943-7006) of ONE fragment with fragment-relative keys row*16+slot
 * (fragment.go:2780-2782).  index/field/view are caller-assigned ids.  Replaces any previous content of
 * (index,field,view,shard).  The library copies; the caller keeps ownership of the bytes. */
int fbgpu_load_fragment(fbgpu_ctx *ctx, uint32_t index, uint32_t field, uint32_t view, uint64_t shard,
                        const uint8_t *roaring, uint64_t nbytes);
/* bulk form (the shape of API.ImportRoaringShard's per-view payloads, api.go:1647; fragment.importRoaringOverwrite
 * fragment.go:2196): n fragments of the same (index,field,view); fragment i is buf[offsets[i], offsets[i+1]) */
int fbgpu_load_fragments(fbgpu_ctx *ctx, uint32_t index, uint32_t field, uint32_t view,
                         const uint64_t *shards, int64_t n, const uint8_t *buf, const uint64_t *offsets);
/* view.deleteFragment (view.go:405): the fragment no longer answers queries; its arena space is reclaimed by fbgpu_compact */
int fbgpu_drop_fragment(fbgpu_ctx *ctx, uint32_t index, uint32_t field, uint32_t view, uint64_t shard);
/* Incremental refresh: what ONE committed write transaction did to ONE fragment, container by container -- the mirror of the
 * Tx.PutContainer / Tx.RemoveContainer calls of the transaction (tx.go:91-96, rbf/tx.go:791-860; the writes reach them through
 * fragment.setBit / importRoaring / ImportRoaringBits, fragment.go:2196, rbf/tx.go:1819).  `roaring` (may be NULL / 0 bytes) holds
 * ONLY the containers that were written, under their fragment-relative keys: each replaces the container stored under its key, or
 * adds it.  removed_keys lists the keys whose containers were deleted.  Containers the transaction did not touch keep their payload
 * where it is in HBM, so the call moves the changed containers' bytes plus the fragment's row / descriptor entries -- not the
 * fragment (fbgpu_load_fragment re-sends it whole); the next fbgpu_commit patches the tables of the touched (view, shard) pairs
 * instead of rebuilding them.  Replaced payloads become holes (fbgpu_stats.dead_bytes) that fbgpu_compact reclaims.  A fragment
 * that is not resident yet is created from the written containers; one whose last container is removed is dropped.  A key that
 * is both written and removed is FBGPU_E_INVALID.  All-or-nothing like the loads. */
int fbgpu_apply_containers(fbgpu_ctx *ctx, uint32_t index, uint32_t field, uint32_t view, uint64_t shard,
                           const uint8_t *roaring, uint64_t nbytes, const uint64_t *removed_keys, int64_t n_removed);
/*  Load the fragments of ONE shard straight from its RBF database, i.e. from the bytes of
 * `<index>/backends/rbf/<shard>/data` (and, if it is not empty, `wal`) -- one RBF DB holds every field/view of a
 * shard (dbshard.go:64-71).  Replaces the per-fragment tx.RoaringBitmap().WriteTo re-serialisation: leaf cells (array /
 * RLE / bitmap page, rbf/rbf.go:489-512) become store containers directly.  names[i] is an RBF bitmap name
 * "~field;view<" (rbfName rbf.go:504, short_txkey/txkey.go:129); fields[i] / views[i] are the ids the caller uses for
 * it in programs.  Names that the file does not hold are skipped (a field without data in this shard); *out_loaded
 * (may be NULL) receives how many were found.  Committed WAL pages override data pages (rbf/tx.go:1269-1273); pages
 * after the WAL's last meta page are ignored.  The library copies; the caller keeps the buffers. */
int fbgpu_load_rbf(fbgpu_ctx *ctx, uint32_t index, uint64_t shard, const uint8_t *data, uint64_t data_bytes,
                   const uint8_t *wal, uint64_t wal_bytes, const char *const *names, const uint32_t *fields,
                   const uint32_t *views, int32_t n_names, int32_t *out_loaded);
/* Same, with the library mapping the files itself (read-only mmap of `<dir>/data` and, when present and non-empty,
 * `<dir>/wal`): the caller hands over the shard's RBF directory instead of reading it into its own memory first.  The
 * mappings are released before the call returns.  FBGPU_E_FORMAT when `data` cannot be opened or mapped. */
int fbgpu_load_rbf_dir(fbgpu_ctx *ctx, uint32_t index, uint64_t shard, const char *dir, const char *const *names,
                       const uint32_t *fields, const uint32_t *views, int32_t n_names, int32_t *out_loaded);
/* pushes pending host-side staging to HBM now (otherwise done lazily by the next query): the point at which loaded
 * fragments become visible to queries, as RBFTx.Commit (rbf.go:189) is for writes in the reference */
int fbgpu_commit(fbgpu_ctx *ctx);
/* Replacing or dropping a fragment leaves its old payload in the arena.  fbgpu_compact() commits what is pending, then
 * copies the live fragments into a fresh arena (device to device) and rebuilds the tables; fbgpu_commit() does the same on
 * its own once at least 256 MiB and half of the arena are dead.  Blocks queries for the duration, like a commit. */
int fbgpu_compact(fbgpu_ctx *ctx);
/* Inspection (FBGPU_DEVICE_NONE contexts only): the container the kernels would find for (index, field, view, shard, row,
 * slot), located by the same resolve() code, with its payload exactly as stored (incl. the padding of the last 16-byte
 * chunk).  *out_type = 0 when the container is absent, else 1 array / 2 bitmap / 3 run. */
int fbgpu_debug_container(fbgpu_ctx *ctx, uint32_t index, uint32_t field, uint32_t view, uint64_t shard, uint64_t row,
                          int32_t slot, uint32_t *out_type, uint32_t *out_card, uint32_t *out_runs,
                          uint8_t *out_payload, uint64_t cap, uint64_t *out_len);

typedef struct {
    uint64_t fragments, containers;
    uint64_t array_containers, bitmap_containers, run_containers;
    uint64_t payload_bytes;   /* roaring payload bytes resident in HBM (array 2n, bitmap 8192, run 4r) */
    uint64_t device_bytes;    /* total HBM held by the store incl. descriptors and padding            */
    uint64_t dead_bytes;      /* arena bytes of replaced / dropped fragments and containers, reclaimed by fbgpu_compact */
    uint64_t full_commits;    /* commits that rebuilt (and re-sent) every lookup table                                */
    uint64_t patch_commits;   /* commits that patched the touched (view, shard) entries and sent only those + the new tails */
} fbgpu_stats;
int fbgpu_get_stats(fbgpu_ctx *ctx, fbgpu_stats *out);

/* ---- bitmap-call programs (replaces executeBitmapCallShard and its children, executor.go:1782-1816) ----
 * A program is the post-order walk of the pql.Call tree of one bitmap call. */
enum {
    FBGPU_OP_ROW = 1,        /* Row(field=row)         executeRowShard executor.go:5120 ; field,view,a=row id        */
    FBGPU_OP_INTERSECT = 2,  /* Intersect(c1..cn)      executeIntersectShard :5357 ; argc=n (0 => FBGPU_E_QUERY)      */
    FBGPU_OP_UNION = 3,      /* Union(c1..cn)          executeUnionShard :5382 ; argc=n (0 => empty row)             */
    FBGPU_OP_DIFFERENCE = 4, /* Difference(c1..cn)     executeDifferenceShard :2950 ; left fold; argc 0 => E_QUERY    */
    FBGPU_OP_XOR = 5,        /* Xor(c1..cn)            executeXorShard :5513 ; left fold; argc 0 => empty row         */
    FBGPU_OP_NOT = 6,        /* Not(c)                 executeNotShard :5554 ; field,view = existence field, a = row  */
    FBGPU_OP_BSI_RANGE = 7,  /* Row(v <op> k)          executeRowBSIGroupShard :5249 -> fragment.rangeOp fragment.go:937
                                field,view = bsig view; a = bitDepth; b = FBGPU_CMP_*; lo(,hi) = base-adjusted
                                predicate(s) exactly as passed to rangeOp/rangeBetween                              */
    FBGPU_OP_EMPTY = 8,      /* empty row (out-of-range BSI predicate, executor.go:5331)                               */
    FBGPU_OP_ALL = 9         /* All(): existence row   executeAllCallShard :5781 ; field,view = existence, a = row    */
};
enum { FBGPU_CMP_EQ = 1, FBGPU_CMP_NEQ = 2, FBGPU_CMP_LT = 3, FBGPU_CMP_LTE = 4, FBGPU_CMP_GT = 5,
       FBGPU_CMP_GTE = 6, FBGPU_CMP_BETWEEN = 7 };

typedef struct {
    uint32_t opcode, field, view, argc;
    uint64_t a, b;
    int64_t lo, hi;
} fbgpu_op; /* 48 bytes */

/* Count(<bitmap call>) over the listed shards (executeCount executor.go:5839-5892: map = per-shard
 * Row.Count(), reduce = u64 add).  *out_total is the sum over this context's shards, all-reduced over
 * the communicator when one is attached.  out_per_shard may be NULL, else receives n_shards counts. */
int fbgpu_count(fbgpu_ctx *ctx, uint32_t index, const fbgpu_op *ops, int32_t n_ops,
                const uint64_t *shards, int64_t n_shards, uint64_t *out_total, uint64_t *out_per_shard);

/* Row.Any() of a bitmap call (row.go:258; roaring.go:4266-4408 intersectionAny lifted to shard granularity): *out_any = 1 as soon
 * as one block of shards holds a column.  Shards are evaluated in blocks of growing size (8, 64, 512, ...), so a non-empty row
 * costs one small launch.  Local to this context: never merged across GPUs (fbgpu_node_any walks the devices itself). */
int fbgpu_any(fbgpu_ctx *ctx, uint32_t index, const fbgpu_op *ops, int32_t n_ops,
              const uint64_t *shards, int64_t n_shards, int32_t *out_any);

/* <bitmap call> returning a Row (mapReduce with Row.Merge, executor.go:1694-1780, row.go:202): writes
 * Pilosa-roaring bytes with absolute keys shard*16+slot and canonical (optimize()) encodings, i.e. what
 * Row.Roaring() (row.go:174-180) yields after Optimize.  If out_cap is too small returns FBGPU_E_NOSPACE
 * and sets *out_len to the needed size.  *out_count receives the row's cardinality (may be NULL). */
int fbgpu_row(fbgpu_ctx *ctx, uint32_t index, const fbgpu_op *ops, int32_t n_ops,
              const uint64_t *shards, int64_t n_shards,
              uint8_t *out_buf, uint64_t out_cap, uint64_t *out_len, uint64_t *out_count);

/* <bitmap call> returning the row's column ids (Row.Columns(), row.go:471, what Extract / Limit / API row responses
 * iterate): ascending absolute column ids (shard * 2^20 + position), expanded on the device from the result bitmaps, so
 * a caller that wants ids does not have to decode roaring containers.  Skips the first `offset` columns and writes at
 * most `limit` (limit < 0: no limit) — executeLimitCall's window.  *out_n = columns written, *out_total = the row's
 * cardinality before the window (may be NULL).  FBGPU_E_NOSPACE when cap is smaller than the window; *out_n then holds
 * the needed capacity. */
int fbgpu_columns(fbgpu_ctx *ctx, uint32_t index, const fbgpu_op *ops, int32_t n_ops,
                  const uint64_t *shards, int64_t n_shards, uint64_t offset, int64_t limit,
                  uint64_t *out_cols, uint64_t cap, uint64_t *out_n, uint64_t *out_total);

/* Values of an int field for the columns of a row: the bulk form of fragment.value (fragment.go:585-617) that Extract
 * (executor.go executeExtract) and Distinct on int fields (executeDistinctShardBSI :2034) are built on.  The row is
 * <filter program> ∩ not-null(field) (n_ops == 0: every column that has a value); for its columns, in ascending order and
 * inside the offset / limit window, out_cols[i] receives the column id and out_vals[i] the stored sign-magnitude value as
 * an int64, i.e. value - bsiGroup.Base (the caller adds Base, field.go:1640).  `view` is the field's bsig_ view,
 * bit_depth <= 63 its current depth.  Same capacity contract as fbgpu_columns. */
int fbgpu_extract(fbgpu_ctx *ctx, uint32_t index, const fbgpu_op *ops, int32_t n_ops, uint32_t field, uint32_t view, int32_t bit_depth,
                  const uint64_t *shards, int64_t n_shards, uint64_t offset, int64_t limit,
                  uint64_t *out_cols, int64_t *out_vals, uint64_t cap, uint64_t *out_n, uint64_t *out_total);

/* Min / Max of an int field over a row (executeMin :1225 / executeMax :1261, fragment.min / max fragment.go:752-838): the row
 * is <filter program> ∩ not-null(field) (n_ops == 0: every column with a value).  *out_val receives the extreme stored
 * value, i.e. value - bsiGroup.Base (the caller adds Base), *out_count how many columns hold it — the reference's ValCount;
 * *out_count == 0 when the row is empty.  One evaluation of the row plus one pass over the bit planes, every plane container
 * read once (the composition from fbgpu_count calls re-reads the planes it has kept).  Not reduced over the communicator:
 * the caller merges per-node ValCounts as it does today (ValCount.Smaller / Larger executor.go:8446-8560). */
int fbgpu_bsi_minmax(fbgpu_ctx *ctx, uint32_t index, const fbgpu_op *ops, int32_t n_ops, uint32_t field, uint32_t view, int32_t bit_depth,
                     const uint64_t *shards, int64_t n_shards, int32_t want_max, int64_t *out_val, uint64_t *out_count);

/* Sum of an int field over a row (executeSum :1119, fragment.sum fragment.go:722-750): *out_count = |<filter> ∩ not-null|,
 * *out_sum = Σ (stored value) over those columns in wrapping int64 arithmetic, i.e. Σ (value - Base); the caller adds
 * count * Base (executeSumCountShard :2203-2206).  One evaluation of the row, one pass over the planes.  Per-node result,
 * like fbgpu_bsi_minmax (ValCount.Add merges nodes). */
int fbgpu_bsi_sum(fbgpu_ctx *ctx, uint32_t index, const fbgpu_op *ops, int32_t n_ops, uint32_t field, uint32_t view, int32_t bit_depth,
                  const uint64_t *shards, int64_t n_shards, int64_t *out_sum, uint64_t *out_count);

/* Per-row counts of one field, optionally intersected with a filter program: the exact part of TopN
 * (fragment.top with explicit ids, fragment.go:1317-1437) and TopK (doTopK executor.go:2705-2746).
 * row_ids != NULL: counts for exactly those rows (out_counts[i] for row_ids[i]).
 * row_ids == NULL: all rows present in the field over the shards; writes every (row id, count) pair with
 * count > 0 sorted by (count desc, row id asc) — the reference's tie order is unspecified (cache.go:464-482).
 * *out_n receives the number of such rows; when it exceeds cap nothing is written and the call returns
 * FBGPU_E_NOSPACE (never a truncated list: Rows() and the TopN candidate set must be complete) — call again
 * with buffers of *out_n entries.
 * Counts are all-reduced over the communicator in the row_ids form. */
int fbgpu_row_counts(fbgpu_ctx *ctx, uint32_t index, uint32_t field, uint32_t view,
                     const uint64_t *row_ids, int32_t n_rows,
                     const fbgpu_op *filter, int32_t n_filter_ops,
                     const uint64_t *shards, int64_t n_shards,
                     uint64_t *out_row_ids, uint64_t *out_counts, int32_t cap, int32_t *out_n);
/* The same counts kept apart per shard, for explicit rows: out_counts[s * n_rows + i] = |Row(row_ids[i]) [∩ filter]| in
 * shards[s] (a shard without the fragment gives zeros).  fragment.top applies its MinThreshold / Tanimoto cut-offs to each
 * shard's own counts before Pairs.Add sums them (fragment.go:1329-1388, executeTopNShards executor.go:2831-2866), so TopN
 * with threshold= / tanimotoThreshold= needs this matrix, not the reduced vector.  Never all-reduced: a shard belongs to
 * one rank. */
int fbgpu_row_counts_per_shard(fbgpu_ctx *ctx, uint32_t index, uint32_t field, uint32_t view,
                               const uint64_t *row_ids, int32_t n_rows, const fbgpu_op *filter, int32_t n_filter_ops,
                               const uint64_t *shards, int64_t n_shards, uint64_t *out_counts);

/* Many fused Intersect+Count pairs in ONE launch: out_counts[i] = |Row(field_a = rows_a[i]) ∩ Row(field_b = rows_b[i])| over
 * the shards — the inner loop of fragment.top with a plain-row Src (count = Src.intersectionCount(row) per candidate
 * row, fragment.go:1367-1372,1416-1420), of the GroupBy leaf (executor.go:8893) and of BenchmarkFragment_IntersectionCount
 * (fragment_internal_test.go:1461), without materialising anything.  All-reduced over the communicator. */
int fbgpu_count_pairs(fbgpu_ctx *ctx, uint32_t index, uint32_t field_a, uint32_t view_a, const uint64_t *rows_a,
                      uint32_t field_b, uint32_t view_b, const uint64_t *rows_b, int32_t n_pairs,
                      const uint64_t *shards, int64_t n_shards, uint64_t *out_counts);

/* Which of the nine container-pair kernels Count(Intersect(Row a, Row b)) exercises and how often: out_hist[4 * ta + tb] = number
 * of (shard, slot) units whose a / b containers have types ta / tb (0 absent, 1 array, 2 bitmap, 3 run).  The reference keeps
 * the same information as statsHit("intersectionCount/...") counters (roaring.go:4477-4614).  Computed on the device. */
int fbgpu_pair_types(fbgpu_ctx *ctx, uint32_t index, uint32_t field_a, uint32_t view_a, uint64_t row_a,
                     uint32_t field_b, uint32_t view_b, uint64_t row_b, const uint64_t *shards, int64_t n_shards, uint64_t out_hist[16]);

/* GroupBy(Rows(f1), Rows(f2), ..., filter=...) with Count aggregate (executeGroupBy executor.go:3176,
 * executeGroupByShard :3918, groupByIterator :8617-8934; reduce = mergeGroupCounts :3728).  row_ids_flat holds
 * the per-field row-id lists concatenated (flat on purpose: cgo forbids nested Go pointers).  out_counts is
 * the dense count tensor, row-major, rightmost field fastest; the caller emits groups with Count>0 in
 * lexicographic order (executor.go:3960).  A shard lacking a fragment for any field contributes nothing
 * (executor.go:8769-8772).  All-reduced over the communicator. */
int fbgpu_groupby(fbgpu_ctx *ctx, uint32_t index, const uint32_t *fields, const uint32_t *views, int32_t n_fields,
                  const uint64_t *row_ids_flat, const int32_t *n_rows,
                  const fbgpu_op *filter, int32_t n_filter_ops,
                  const uint64_t *shards, int64_t n_shards, uint64_t *out_counts);

/* ---- multi-GPU reduce (replaces the HTTP fan-in of mapReduce/remoteExec, executor.go:6392-6533) ----
 * One context (process) per GPU; rank 0 creates the id, every rank joins.  When a communicator is attached,
 * count / row_counts / groupby results are summed with one ncclAllReduce(uint64,sum) on the device before
 * the single D2H copy.  NCCL is resolved at run time (dlopen libnccl.so.2). */
#define FBGPU_NCCL_ID_BYTES 128
int fbgpu_comm_unique_id(uint8_t id[FBGPU_NCCL_ID_BYTES]);
int fbgpu_comm_init(fbgpu_ctx *ctx, int32_t n_ranks, int32_t rank, const uint8_t id[FBGPU_NCCL_ID_BYTES]);
int fbgpu_comm_destroy(fbgpu_ctx *ctx);
/* Fused Count merge over NVLink peer memory (optional, Count only): every rank exports a small mailbox through CUDA IPC
 * (fbgpu_comm_p2p_handle), the host exchanges the 64-byte handles (any transport), every rank maps its peers
 * (fbgpu_comm_p2p_open).  From then on fbgpu_count() sums the per-GPU totals inside the counting kernel itself (last CTA
 * stores to all peers' mailboxes, waits for theirs) instead of launching a separate all-reduce. */
int fbgpu_comm_p2p_handle(fbgpu_ctx *ctx, uint8_t out_handle[64]);
int fbgpu_comm_p2p_open(fbgpu_ctx *ctx, int32_t n_ranks, int32_t rank, const uint8_t *handles /* n_ranks x 64 bytes */);
int fbgpu_comm_p2p_disable(fbgpu_ctx *ctx);   /* fall back to the NCCL merge (e.g. when a peer could not be mapped) */

/* In-process form of the same exchange: the contexts of ONE process (one per GPU, one caller thread each) are wired to each
 * other through peer access instead of CUDA IPC.  ctxs[r] becomes rank r. */
int fbgpu_comm_p2p_open_local(fbgpu_ctx *const *ctxs, int32_t n_ranks);
/* The wait for a peer's count inside the kernel is bounded (FBGPU_P2P_TIMEOUT_MS, default 2000): a peer that died, or ranks
 * that issued their collective queries in different orders, make fbgpu_count() return FBGPU_E_COMM instead of hanging the GPU.
 * The multi-process forms keep NCCL's precondition: every rank issues its collective queries in the same order, from one thread
 * at a time.  A process whose threads query concurrently (FeatureBase: executor.go:6449-6533) uses fbgpu_node below. */

/* ---- every GPU of one process behind one handle (SURVEY §8(b) fbgpu_init(device_ordinals, n); replaces mapReduce's local
 *      fan-out + reduce, executor.go:6449-6533, 6742-6812) ----
 * A node owns one context per listed device.  Shard s lives on device slot (s / shard_block) % n_devices: contiguous blocks of
 * shard_block shards per GPU (shard_block = ceil(total shards / n_devices) gives one range per GPU).  Every
 * fbgpu_node_* query takes the caller's whole shard list, runs each device's share concurrently on that device (own worker
 * thread, stream and workspace per call) and merges the per-device results on the host with the reference's reducers (u64
 * add: Count executor.go:5880, Pairs.Add cache.go:464, mergeGroupCounts executor.go:3728; Row.Merge row.go:202).  Any number
 * of threads may call concurrently; calls never share result buffers, and a failure on one device fails only that call.
 * The same device ordinal may be listed more than once (two contexts on one GPU; used by the tests). */
typedef struct fbgpu_node fbgpu_node;
int fbgpu_node_init(const int32_t *device_ordinals, int32_t n_devices, uint64_t shard_block, fbgpu_node **out);
void fbgpu_node_shutdown(fbgpu_node *node);
int32_t fbgpu_node_devices(const fbgpu_node *node);
int32_t fbgpu_node_owner(const fbgpu_node *node, uint64_t shard);      /* device slot that holds the shard */
fbgpu_ctx *fbgpu_node_ctx(fbgpu_node *node, int32_t slot);             /* the slot's context (stats, counters); owned by the node */
int fbgpu_node_load_fragment(fbgpu_node *node, uint32_t index, uint32_t field, uint32_t view, uint64_t shard,
                             const uint8_t *roaring, uint64_t nbytes);
int fbgpu_node_load_fragments(fbgpu_node *node, uint32_t index, uint32_t field, uint32_t view,
                              const uint64_t *shards, int64_t n, const uint8_t *buf, const uint64_t *offsets);
int fbgpu_node_load_rbf_dir(fbgpu_node *node, uint32_t index, uint64_t shard, const char *dir, const char *const *names,
                            const uint32_t *fields, const uint32_t *views, int32_t n_names, int32_t *out_loaded);
int fbgpu_node_drop_fragment(fbgpu_node *node, uint32_t index, uint32_t field, uint32_t view, uint64_t shard);
int fbgpu_node_apply_containers(fbgpu_node *node, uint32_t index, uint32_t field, uint32_t view, uint64_t shard,
                                const uint8_t *roaring, uint64_t nbytes, const uint64_t *removed_keys, int64_t n_removed);
int fbgpu_node_commit(fbgpu_node *node);
int fbgpu_node_get_stats(fbgpu_node *node, fbgpu_stats *out);           /* summed over the devices */
/* same contracts as the fbgpu_* calls of the same name, over all devices */
int fbgpu_node_count(fbgpu_node *node, uint32_t index, const fbgpu_op *ops, int32_t n_ops,
                     const uint64_t *shards, int64_t n_shards, uint64_t *out_total, uint64_t *out_per_shard);
int fbgpu_node_any(fbgpu_node *node, uint32_t index, const fbgpu_op *ops, int32_t n_ops,
                   const uint64_t *shards, int64_t n_shards, int32_t *out_any);
int fbgpu_node_row(fbgpu_node *node, uint32_t index, const fbgpu_op *ops, int32_t n_ops,
                   const uint64_t *shards, int64_t n_shards, uint8_t *out_buf, uint64_t out_cap, uint64_t *out_len, uint64_t *out_count);
int fbgpu_node_count_pairs(fbgpu_node *node, uint32_t index, uint32_t field_a, uint32_t view_a, const uint64_t *rows_a,
                           uint32_t field_b, uint32_t view_b, const uint64_t *rows_b, int32_t n_pairs,
                           const uint64_t *shards, int64_t n_shards, uint64_t *out_counts);
int fbgpu_node_row_counts(fbgpu_node *node, uint32_t index, uint32_t field, uint32_t view, const uint64_t *row_ids, int32_t n_rows,
                          const fbgpu_op *filter, int32_t n_filter_ops, const uint64_t *shards, int64_t n_shards, uint64_t *out_counts);
int fbgpu_node_groupby(fbgpu_node *node, uint32_t index, const uint32_t *fields, const uint32_t *views, int32_t n_fields,
                       const uint64_t *row_ids_flat, const int32_t *n_rows, const fbgpu_op *filter, int32_t n_filter_ops,
                       const uint64_t *shards, int64_t n_shards, uint64_t *out_counts);
int fbgpu_node_bsi_sum(fbgpu_node *node, uint32_t index, const fbgpu_op *ops, int32_t n_ops, uint32_t field, uint32_t view, int32_t bit_depth,
                       const uint64_t *shards, int64_t n_shards, int64_t *out_sum, uint64_t *out_count);
int fbgpu_node_bsi_minmax(fbgpu_node *node, uint32_t index, const fbgpu_op *ops, int32_t n_ops, uint32_t field, uint32_t view, int32_t bit_depth,
                          const uint64_t *shards, int64_t n_shards, int32_t want_max, int64_t *out_val, uint64_t *out_count);

/* Inspection (any context): the stack-machine program the library would run for `ops` -- records of 16 bytes {u8 op, u8 pad[3],
 * u32 view slot, u64 row} (csrc/fbgpu_types.h DevOp); *out_depth = operand stack depth.  With index == 0xffffffff,
 * fbgpu_debug_container() takes such a view slot in `field`. */
int fbgpu_debug_compile(fbgpu_ctx *ctx, uint32_t index, const fbgpu_op *ops, int32_t n_ops, uint8_t *out, int32_t cap_ops,
                        int32_t *out_n, int32_t *out_depth);

/* ---- instrumentation (the counters the reference keeps under the roaringstats tag, statsHit()) ---- */
typedef struct {
    uint64_t kernel_launches;   /* kernels of this library launched since init       */
    uint64_t queries;
    float last_query_gpu_ms;    /* CUDA-event time of the last query's kernels        */
    uint32_t pair_kernel_queries; /* Count(Intersect(Row, Row)) queries answered by the fused pair_count_kernel (low 32 bits) */
    uint64_t last_algo_bytes;   /* algorithmic bytes of the last query   */
    uint64_t groupby_units;     /* (shard, slot) units of GroupBy queries so far ...                          */
    uint64_t groupby_fallback_units; /* ... and how many of them the warp-per-unit kernel handed to the CTA kernel */
} fbgpu_counters;
int fbgpu_get_counters(fbgpu_ctx *ctx, fbgpu_counters *out);
/* algorithmic-bytes accounting: roaring payload bytes and container count of the given rows (NULL = all
 * rows) of a field over the shards, from the host-side directory */
int fbgpu_rows_payload_bytes(fbgpu_ctx *ctx, uint32_t index, uint32_t field, uint32_t view, const uint64_t *row_ids, int32_t n_rows,
                             const uint64_t *shards, int64_t n_shards, uint64_t *out_payload, uint64_t *out_containers);
/* the CUDA stream queries of the calling thread run on (cudaStream_t), for external event timing */
void *fbgpu_stream(fbgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
