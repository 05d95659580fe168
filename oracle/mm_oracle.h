/*
 * mm_oracle.h — C API of the CPU ORACLE for the ModelMesh placement / LRU hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a line-by-line CPU restatement of the reference's
 * Java (see mm_oracle.cpp header for the file:line map).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (modelmesh_b200/csrc, include/mmplace.h) never links, imports or calls anything in oracle/.
 *
 * Parity status: the reference's own tests hold no unit vectors for this path;
 * the integration scenarios they do hold (EvictionsModelMeshTest, ModelMeshEvictionsTest,
 * ModelMeshErrorPropagationTest) are restated as known-answer tests in tests/test_oracle_golden.py.
 * The Java reference cannot be executed in this image (no JDK, un-vendored jars), so everything
 * those scenarios do not pin is "parity unpinned by reference tests" — pinned only by the
 * line-by-line correspondence documented in mm_oracle.cpp.
 */
#ifndef MM_ORACLE_H
#define MM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_fleet orc_fleet;

/* Numeric part of InstanceRecord (InstanceRecord.java:37-73). */
typedef struct {
  int64_t lru_time;   /* Long.MAX_VALUE when empty */
  int64_t capacity;   /* units of 8 KiB */
  int64_t used;
  int64_t start_time;
  int64_t vers;
  int32_t count;
  int32_t l_threads;
  int32_t l_in_prog;
  int32_t rpm;
  int32_t shutting_down;
  int32_t active;     /* present in litelinks' service-instance map (siMap, MM:4778) */
} orc_inst_t;

/* ModelMesh.ClusterStats (MM:1570-1591) */
typedef struct {
  int64_t total_capacity, total_free, global_lru;
  int32_t instance_count, model_copy_count;
} orc_stats_t;

/* One placement decision = one call of CacheMissForwardingLB.getNext (MM:4776-5004). */
typedef struct {
  int32_t type_idx;     /* index into the type-name table passed to orc_get_next_batch; -1 = a type with no config */
  int32_t self;         /* instance idx of the caller ("instanceId") */
  int32_t fresh_idx;    /* index into fresh[] (getFreshInstanceRecord, MM:5369), -1 = use self's published row with rpm forced to 0 */
  int32_t favour_self;  /* CacheMissExcludeSet.favourSelf */
  int64_t last_used;    /* CacheMissExcludeSet.lastUsedTime */
  uint64_t decision_id; /* feeds the replacement of ThreadLocalRandom (note N4) */
} orc_decision_t;

enum { ORC_NONE = -1, ORC_SELF = -2 };

typedef struct {
  int32_t target;        /* instance idx, ORC_NONE (null) or ORC_SELF (ABORT_REQUEST) */
  int32_t n_candidates;  /* candidates.size() at MM:4939 (0 if returned earlier) */
  int32_t n_remaining;   /* remainingCount after the rpm filter */
  int32_t pick_index;    /* index-th non-null candidate */
  int32_t best;          /* instance idx of bestIid after preferred handling, -1 if none */
  int32_t flags;         /* bit0: replicaset filter retried (MM:4798); bit1: simpleCase at MM:4889;
                            bit2: bestIsFull; bit3: returned via favourSelf short-circuit */
} orc_result_t;

orc_fleet *orc_create(int64_t min_space_units, int64_t min_churn_age_ms, int32_t default_model_size_units);
void orc_destroy(orc_fleet *);

/* handleInstanceTableChange (MM:1455-1568).  type: 0 ENTRY_ADDED, 1 ENTRY_UPDATED, 2 ENTRY_DELETED.
 * idx is the caller's dense handle for instance id `id` (a bijection).  loc/zone may be NULL.
 * Strings are UTF-8; they are compared as UTF-16 code units like java.lang.String.compareTo. */
int orc_instance_event(orc_fleet *, int type, int32_t idx, const orc_inst_t *rec, const char *id,
                       const char *loc, const char *zone, const char *const *labels, int32_t n_labels,
                       int64_t now_ms);
/* Test-harness shortcut for big fleets: same final state as n ENTRY_ADDED events (instance idx = position) on an empty
 * fleet followed by orc_tc_converge(), without the per-event O(N) work (see Fleet::bulkAdd). locs/zones entries may be NULL. */
int orc_bulk_add(orc_fleet *, int32_t n, const orc_inst_t *recs, const char *const *ids, const char *const *locs,
                 const char *const *zones, const int32_t *label_off, const char *const *labels);
/* litelinks siMap membership (MM:4778) for an instance that is in the table */
int orc_set_active(orc_fleet *, int32_t idx, int32_t active);

/* TypeConstraintManager.typeMappingsUpdated (TCM:607-668).  The config is passed pre-parsed:
 * n types; for type t, names[t], then req_off[t]..req_off[t+1] index into req_labels, same for pref.
 * Passing n = -1 disables type constraints (typeConstraints == null). */
int orc_types_set(orc_fleet *, int32_t n, const char *const *names, const int32_t *req_off,
                  const char *const *req_labels, const int32_t *pref_off, const char *const *pref_labels);

/* Runs TypeConstraintManager.refreshPerTypeInstanceSets (TCM:680-725) once more on the current map with every
 * instance already in clusterState.  The reference reaches this state whenever a refresh happens after the fleet has
 * stopped changing; calling it explicitly removes the arrival-order lag of quirk N10 so that a snapshot-based solver
 * can be compared without replaying the exact event order.  No code other than the literal restatement runs. */
int orc_tc_converge(orc_fleet *);
/* Bulk-load switch: while set, TCM.instanceAdded (TCM:513-526) skips its refreshPerTypeInstanceSets call (an
 * O(instances x types) pass per event, quadratic for a 10k-instance fleet).  The sets it would have produced are a
 * pure function of the membership sets, so a following orc_tc_converge() yields the same state. */
int orc_tc_defer_refresh(orc_fleet *, int defer);

/* Directly set UpgradeTracker.likelyReplacedReplicaSets keys (UT:71); the event-driven tracker
 * (UT:120-187) also runs inside orc_instance_event. */
int orc_set_replaced_replicasets(orc_fleet *, const char *const *prefixes, int32_t n);
int orc_get_replaced_replicasets(orc_fleet *, char *buf, int32_t cap); /* comma-joined, returns count */

/* Read back TCM state for a type name: fills allowed/preferred membership (0/1 per instance idx < n_idx);
 * *allowed_null / *preferred_null tell whether the Java Set is null. */
int orc_type_sets(orc_fleet *, const char *type, int32_t n_idx, uint8_t *allowed, int32_t *allowed_null,
                  uint8_t *preferred, int32_t *preferred_null);

/* Sorted fleet (clusterState iteration order under PLACEMENT_ORDER, MM:4646-4703): writes instance idx. */
int orc_cluster_order(orc_fleet *, int32_t *out_idx, int32_t cap);
/* PLACEMENT_ORDER.compare on two instances currently in clusterState */
int orc_compare(orc_fleet *, int32_t idx1, int32_t idx2);

/* clusterStats (global) and per-PTS partition stats in TCM.getPartitionStats order (TCM:264-292).
 * part_pts_ids returns an opaque id per partition also reported by orc_instance_partition. */
int orc_cluster_stats(orc_fleet *, orc_stats_t *out);
int orc_partition_stats(orc_fleet *, orc_stats_t *out, int32_t *part_ids, int32_t cap);
int orc_instance_partition(orc_fleet *, int32_t idx);
int orc_type_stats(orc_fleet *, const char *type, orc_stats_t *out); /* ModelMesh.typeSetStats MM:1432-1438 */

/* Batch of getNext calls against the current state.
 * excl_off[n+1]/excl_idx: per decision the union of CacheMissExcludeSet members (tried ∪ loaded ∪ failed ∪ explicit).
 * Optional outputs (may be NULL): cand_off[n+1] + cand_idx/cand_load/cand_keep[cap] = ordered candidates,
 * instReqLoad and whether each survived the rpm filter. Returns number of candidate slots needed or <0. */
int64_t orc_get_next_batch(orc_fleet *, int32_t n, const orc_decision_t *dec, const char *const *type_names,
                           int32_t n_types, const orc_inst_t *fresh, int32_t n_fresh, const int64_t *excl_off,
                           const int32_t *excl_idx, int64_t now_ms, uint64_t seed, int32_t threads,
                           orc_result_t *out, int64_t *cand_off, int32_t *cand_idx, int32_t *cand_load,
                           uint8_t *cand_keep, int64_t cand_cap);

/* the same decisions with the entries walked through a rank-ordered array built once per batch instead of the ordered set
 * (CPU-baseline mode (ii) "dense" of BASELINE.md §4) */
int64_t orc_get_next_batch_dense(orc_fleet *, int32_t n, const orc_decision_t *dec, const char *const *type_names,
                                 int32_t n_types, const orc_inst_t *fresh, int32_t n_fresh, const int64_t *excl_off,
                                 const int32_t *excl_idx, int64_t now_ms, uint64_t seed, int32_t threads, orc_result_t *out);

/* ---- time-ordered weighted LRU (clhm/ConcurrentLinkedHashMap + LinkedDeque) ---- */
typedef struct orc_lru orc_lru;
typedef struct {
  int32_t op;       /* 0 INSERT (putIfAbsent k,v,lastUsed) 1 TOUCH (get k,lastUsed) 2 RESIZE (replace weight, quiet)
                       3 REMOVE 4 SET_CAPACITY (weight = new capacity) 5 FORCE_TIME (forceSetLastUsedTime) */
  int32_t key;
  int64_t weight;
  int64_t last_used; /* 0 = now */
} orc_lru_event_t;
typedef struct { int32_t key; int32_t event; int64_t last_used; int64_t weight; } orc_eviction_t;
orc_lru *orc_lru_create(int64_t capacity);
void orc_lru_destroy(orc_lru *);
/* applies events in order; evictions appended in listener order. returns number of evictions (may exceed cap) */
int64_t orc_lru_apply(orc_lru *, const orc_lru_event_t *ev, int64_t n, int64_t now_ms, orc_eviction_t *out, int64_t cap);
int64_t orc_lru_oldest_time(orc_lru *);
int64_t orc_lru_weighted_size(orc_lru *);
int64_t orc_lru_size(orc_lru *);
/* ascending (oldest first) dump: keys / lastUsed / weights */
int64_t orc_lru_dump(orc_lru *, int32_t *keys, int64_t *last_used, int64_t *weights, int64_t cap);

/* ---- capacity constants (MM:749-755, 767-769) ---- */
int64_t orc_unload_reserve_units(int64_t cache_capacity_units, int32_t loading_threads, int32_t default_model_size_units);
int64_t orc_min_space_units(int64_t cap_units, int32_t loading_threads, int32_t default_model_size_units, int32_t has_unload_manager);

/* ---- loadLocal admission rules (MM:5145-5148, 5185-5197) and churn guard (MM:3872-3884) ---- */
/* returns 1 if the load is rejected by the churn guard */
int orc_churn_reject(int64_t capacity, int64_t weighted_size, int64_t oldest_time, int64_t min_space_units,
                     int64_t min_churn_age_ms, int64_t now_ms);
/* returns 1 if loadLocal aborts early at MM:5187 */
int orc_early_reject(int64_t abs_size, int64_t capacity, int64_t weighted_size, int64_t oldest_time, int64_t last_used);

/* ---- reaper proactive-load selection (MM:6574-6577, 6616-6735) over the current fleet state ---- */
typedef struct {
  int64_t last_used;
  int32_t type_idx;    /* index into type_names, -1 none */
  int32_t n_loaded;    /* insts.size() */
  int32_t n_failed;    /* failInsts.size() */
  int32_t pad;
} orc_model_t;
/* Runs candidate collection over all n models then triggerProactiveLoadsForInstanceSubset for the
 * partition `part_id` (-1 = no type constraints / whole cluster).  Writes selected model idx MRU-first.
 * `taken` (n bytes, in/out) mirrors allCandidates.set(i,null). */
int64_t orc_reaper_select(orc_fleet *, int32_t n, const orc_model_t *models, const char *const *type_names,
                          int32_t n_types, int32_t part_id, int64_t now_ms, uint8_t *taken, int32_t *out_models,
                          int64_t cap);

uint64_t orc_hash64(uint64_t seed, uint64_t decision_id);

/* ---- closed loop (mm_sim.inc): placement -> admission -> LRU insert -> eviction -> rebalance -> republish, one epoch
 * (= one republish window) per call; see the header of mm_sim.inc for the epoch semantics both sides implement ---- */
typedef struct orc_sim orc_sim;
typedef struct { int64_t last_used; int32_t type_idx; int32_t size_units; } orc_sim_model_t;
typedef struct { int32_t type; int32_t model; int32_t caller; uint32_t u; int64_t t; } orc_sim_event_t; /* type 0 REQUEST, 1 REMOVE */
typedef struct { int32_t model, self, target, n_candidates, status, event; } orc_sim_decision_t;
/* status: 0 accepted, 1 nowhere to load (getNext null), 2 churn guard (MM:3872-3884), 3 placeholder evicted immediately
 * (MM:5145-5148), 4 early reject (MM:5185-5190), 5 evicted while growing (MM:2102-2106), 6 entry already there, 7 follow-on
 * skipped (model has a copy again), 8 invalid, 9 accepted but evicted again later in the same epoch.
 * event: index of the REQUEST that caused it, or -1 - k for the k-th queued ensureLoadedElsewhere of the previous epoch */
typedef struct { int32_t instance, model; int64_t last_used; int32_t weight, order, reload; } orc_sim_eviction_t;
orc_sim *orc_sim_create(orc_fleet *, int32_t n_models, const orc_sim_model_t *models, const char *const *type_names,
                        int32_t n_types, const int64_t *edge_off, const int32_t *edge_inst, const int32_t *n_loaded,
                        int32_t n_instances, const int64_t *capacity, int64_t load_timeout_ms, int64_t last_published_ms);
void orc_sim_destroy(orc_sim *);
int orc_sim_seed(orc_sim *, int32_t instance, int32_t n, const int32_t *model, const int64_t *last_used, const int32_t *weight,
                 const int64_t *load_ts, int64_t now_ms);
/* returns the number of ensureLoadedElsewhere requests queued for the next epoch */
int64_t orc_sim_step(orc_sim *, const orc_sim_event_t *ev, int32_t n, int64_t now0, int64_t now1, uint64_t seed,
                     orc_sim_decision_t *dec_out, int32_t dec_cap, int32_t *n_dec_out, orc_sim_eviction_t *evict_out,
                     int32_t evict_cap, int32_t *n_evict_out, orc_inst_t *rows_out, int32_t *published_out);
int64_t orc_sim_model_copies(orc_sim *, int32_t model, int32_t *out, int32_t cap, int64_t *last_used);
int64_t orc_sim_lru_state(orc_sim *, int32_t instance, int64_t *oldest, int64_t *weighted, int64_t *count);
int64_t orc_sim_coalesced(orc_sim *);

/* ---- a14: scale-up / scale-down arithmetic (MM:5640-5806, 5835-5870, 6197-6335) ---- */
int orc_second_copy_trigger(int32_t *i1, int32_t *i2, int32_t iteration, int32_t min_age_iters, int32_t max_age_iters,
                            int64_t total_free, int64_t total_capacity, int64_t global_lru, int64_t now,
                            int64_t second_copy_lru_threshold_ms);
int32_t orc_scaleup_copies(int64_t count, int64_t time_delta_ms, int32_t scale_up_rpms, int32_t loaded_count, int32_t failed_count,
                           int32_t suitable_inst_count, int32_t excluded_count, int32_t excluded_not_holding, int32_t recently_loaded,
                           int32_t *rpm_out);
int32_t orc_scaleup_exclude_set(orc_fleet *, int32_t self, int32_t scale_up_rpms, int32_t our_rpm, uint8_t *marks, int32_t n_idx);
int orc_loaded_since(const int32_t *inst, const int64_t *load_ts, int32_t n, int64_t cutoff, int32_t ignore_instance);
typedef struct { int32_t instance, model; int64_t count, last_used, last_heavy; int32_t i1, i2, weight, flags; } orc_scale_in_t;
typedef struct {
  int64_t now, last_check_time;
  int32_t iteration, scale_up_rpm_threshold, second_copy_min_age_iters, second_copy_max_age_iters;
  int64_t second_copy_lru_threshold_ms, rate_check_interval_ms, assume_completed_ms, second_copy_remove_max_age_ms;
  int32_t can_remove, pad;
} orc_scale_params_t;
typedef struct { int32_t action, copies_to_load; int64_t load_last_used; int32_t rpm, i1, i2, set_heavy, remove; } orc_scale_out_t;
int orc_rate_task_eval(orc_fleet *, int32_t n, const orc_scale_in_t *in, const orc_scale_params_t *p, const char *const *type_names,
                       int32_t n_types, const int32_t *type_idx, const int64_t *edge_off, const int32_t *edge_inst, const int64_t *edge_ts,
                       const int32_t *n_loaded, orc_scale_out_t *out);
int orc_janitor_eval(orc_fleet *, int32_t n, const orc_scale_in_t *in, const orc_scale_params_t *p, const int64_t *edge_off,
                     const int32_t *edge_inst, const int64_t *edge_ts, const int32_t *n_loaded, const int64_t *last_unload_time,
                     orc_scale_out_t *out);
int32_t orc_prune_missing(orc_fleet *, int32_t self, const int32_t *inst, const int64_t *ts, int32_t n, int64_t now, int64_t assume_gone_ms,
                          int64_t *missing_since, uint8_t *pruned);
int orc_scale_down(orc_fleet *, int32_t self, const int32_t *copies, const int64_t *load_ts, int32_t n_copies, int64_t last_used,
                   int64_t now, int64_t last_heavy_time, int64_t last_unload_time, int64_t last_check_time, int64_t interval_count,
                   int32_t scale_up_rpm_threshold, int64_t rate_check_interval_ms, int64_t second_copy_remove_max_age_ms,
                   int32_t local_stats_known);

#ifdef __cplusplus
}
#endif
#endif
