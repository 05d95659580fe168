// Stand-alone driver of the C ABI (no Python, no torch) for compute-sanitizer runs of the scoring kernels:
//   compute-sanitizer --tool memcheck|racecheck|synccheck tools/sanitize/driver [n_instances n_models n_decisions]
// A small random fleet with type constraints, one commit, one traced batch (k_place, cooperative) and one untraced
// batch (k_place_lanes) whose results must agree.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/mmplace.h"

static uint64_t rs = 88172645463325252ULL;
static uint32_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }

int main(int argc, char **argv) {
  const int NI = argc > 1 ? atoi(argv[1]) : 1500, NM = argc > 2 ? atoi(argv[2]) : 4000, ND = argc > 3 ? atoi(argv[3]) : 6000;
  mmp_config cfg = {2560, 600000, 2560, NI, NM, 0, 0, 1, 0, 0};
  mmp_fleet *f = nullptr;
  if (mmp_fleet_create(&cfg, &f) < 0) { fprintf(stderr, "create: %s\n", mmp_last_error(nullptr)); return 2; }
  mmp_types_set_json(f, "{\"ta\":{\"required\":[\"l1\"]},\"tb\":{\"preferred\":[\"l2\",\"l3\"]},\"tc\":{\"required\":[\"l2\"],\"preferred\":[\"l1\"]}}");
  const int ta = mmp_type_id(f, "ta"), tb = mmp_type_id(f, "tb"), tc = mmp_type_id(f, "tc");
  const int64_t now = 1760000000000LL;
  for (int i = 0; i < NI; i++) {
    mmp_instance_row r = {};
    r.capacity = 25600 + (rnd() % 4) * 1000;
    r.used = rnd() % 3 == 0 ? r.capacity - (rnd() % 3000) : rnd() % (uint32_t)r.capacity;
    r.lru_time = now - (int64_t)(rnd() % 7200000);
    r.count = rnd() % 30; r.l_threads = 8; r.l_in_prog = rnd() % 3; r.rpm = rnd() % 400; r.start_time = now - 86400000; r.vers = 1;
    r.active = 1;
    std::string id = "pod-" + std::to_string(100000 + i);
    const char *labs[3]; int nl = 0;
    if (rnd() % 2) labs[nl++] = "l1";
    if (rnd() % 3 == 0) labs[nl++] = "l2";
    if (rnd() % 4 == 0) labs[nl++] = "l3";
    if (mmp_instance_upsert(f, i, &r, id.c_str(), nullptr, (i % 3) ? "z1" : "z2", labs, nl) < 0) { fprintf(stderr, "upsert: %s\n", mmp_last_error(f)); return 2; }
  }
  for (int m = 0; m < NM; m++) {
    mmp_model_row r = {};
    r.last_used = now - (int64_t)(rnd() % 100000000); r.size_units = 256 + rnd() % 4000;
    const int ts[4] = {0, ta, tb, tc};
    r.type_id = (uint16_t)ts[rnd() % 4];
    int32_t ids[6]; int n = rnd() % 7 == 0 ? 6 : rnd() % 4;
    for (int k = 0; k < n; k++) ids[k] = (int32_t)(rnd() % NI);
    r.copy_count = (uint8_t)n;
    if (mmp_model_upsert(f, m, &r, ids, n) < 0) { fprintf(stderr, "model: %s\n", mmp_last_error(f)); return 2; }
  }
  if (mmp_fleet_commit(f) < 0) { fprintf(stderr, "commit: %s\n", mmp_last_error(f)); return 2; }
  std::vector<mmp_decision_in> d(ND);
  std::vector<int32_t> extra;
  for (int i = 0; i < ND; i++) {
    d[i].model = (int32_t)(rnd() % NM); d[i].self = (int32_t)(rnd() % NI); d[i].last_used = now - (int64_t)(rnd() % 4000000);
    d[i].flags = (rnd() % 3 == 0 ? MMP_DF_FAVOUR_SELF : 0) | (rnd() % 2 ? MMP_DF_MODEL_LAST_USED : 0);
    d[i].fresh = -1; d[i].extra_off = (int32_t)extra.size(); d[i].extra_n = rnd() % 9 == 0 ? 2 : 0;
    for (int k = 0; k < d[i].extra_n; k++) extra.push_back((int32_t)(rnd() % NI));
  }
  std::vector<mmp_decision_out> a(ND), b(ND);
  std::vector<mmp_decision_trace> tr(ND);
  if (mmp_place_batch(f, d.data(), ND, nullptr, 0, extra.data(), (int32_t)extra.size(), a.data(), now, 7) < 0) { fprintf(stderr, "place: %s\n", mmp_last_error(f)); return 2; }
  if (mmp_place_batch_trace(f, d.data(), ND, nullptr, 0, extra.data(), (int32_t)extra.size(), b.data(), tr.data(), nullptr, now, 7) < 0) { fprintf(stderr, "trace: %s\n", mmp_last_error(f)); return 2; }
  int bad = 0, none = 0;
  for (int i = 0; i < ND; i++) { bad += a[i].target != b[i].target || a[i].n_candidates != b[i].n_candidates; none += a[i].target == MMP_TARGET_NONE; }
  std::vector<int32_t> self(ND);
  for (int i = 0; i < ND; i++) self[i] = d[i].self;
  std::vector<mmp_decision_out> c(NM < ND ? NM : ND);
  if (mmp_place_sweep(f, 0, (int32_t)c.size(), self.data(), 1, nullptr, c.data(), now, 7) < 0) { fprintf(stderr, "sweep: %s\n", mmp_last_error(f)); return 2; }
  printf("driver: %d decisions, lanes vs traced mismatches %d, none %d, launches %lld\n", ND, bad, none, (long long)mmp_kernel_launches(f));
  mmp_fleet_destroy(f);
  return bad ? 1 : 0;
}
