/* fake_nccl.c -- TEST INFRASTRUCTURE: a stand-in for libnccl.so.2 (LASER_B200_NCCL_LIB) so that the host logic of the
 * row-sharded entry points (laser_b200/csrc/capi_multi.inc) runs in the CPU suite against the host-emulated library.
 * One process, "devices" are just indices; a broadcast registered inside ncclGroupStart / ncclGroupEnd is performed at
 * GroupEnd: the root's buffer is copied into every other rank's buffer (host memory is device memory in the emulation).
 * Outside a group (the one-rank-per-process usage, played sequentially by the test: the root's calls first) every call of the
 * root appends its buffer to the communicator's queue and the i-th call of another rank copies from the i-th entry (a
 * row-sharded product may broadcast B in several pieces). */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct fakeComm { int rank, nranks, group_id; } fakeComm;
typedef struct { char internal[128]; } ncclUniqueId;
typedef struct { fakeComm *comm; void *buf; size_t bytes; int root; } PendingOp;

static int g_depth = 0, g_next_group = 1, g_npending = 0, g_calls = 0;
static PendingOp g_pending[64];
#define QLEN 64
static const void *g_root_buf[256][QLEN];   /* by group id: the root's buffers, in call order (sequential mode) */
static size_t g_root_bytes[256][QLEN];
static unsigned g_root_n[256], g_rank_pos[256][16];

int fake_nccl_broadcast_calls(void) { return g_calls; }

int ncclGetUniqueId(ncclUniqueId *id) { memset(id, 0, sizeof *id); id->internal[0] = (char)g_next_group; return 0; }
int ncclCommInitAll(fakeComm **comms, int n, const int *devs) {
  (void)devs;
  for (int i = 0; i < n; ++i) {
    comms[i] = (fakeComm *)malloc(sizeof(fakeComm));
    comms[i]->rank = i; comms[i]->nranks = n; comms[i]->group_id = g_next_group;
  }
  ++g_next_group;
  return 0;
}
int ncclCommInitRank(fakeComm **comm, int n, ncclUniqueId id, int rank) {
  *comm = (fakeComm *)malloc(sizeof(fakeComm));
  (*comm)->rank = rank; (*comm)->nranks = n; (*comm)->group_id = id.internal[0];
  return 0;
}
int ncclCommDestroy(fakeComm *c) { free(c); return 0; }
int ncclGroupStart(void) { ++g_depth; return 0; }
int ncclBroadcast(const void *send, void *recv, size_t count, int dtype, int root, fakeComm *comm, void *stream) {
  (void)stream;
  if (send != recv || dtype != 7 || g_npending >= 64) return 5; /* ncclInvalidUsage */
  if (g_depth == 0) {   /* sequential mode */
    const int g = comm->group_id & 255;
    ++g_calls;
    if (comm->rank == root) {
      g_root_buf[g][g_root_n[g] % QLEN] = recv; g_root_bytes[g][g_root_n[g] % QLEN] = count * 4; ++g_root_n[g];
      return 0;
    }
    if (comm->rank < 0 || comm->rank >= 16) return 5;
    /* a rank that skipped earlier sequences (it joined later) starts at the oldest entry still queued for this size */
    unsigned *pos = &g_rank_pos[g][comm->rank];
    if (g_root_n[g] - *pos > QLEN) *pos = g_root_n[g] - QLEN;
    while (*pos < g_root_n[g] && g_root_bytes[g][*pos % QLEN] != count * 4) ++*pos;
    if (*pos >= g_root_n[g]) return 5;
    memcpy(recv, g_root_buf[g][*pos % QLEN], count * 4);
    ++*pos;
    return 0;
  }
  g_pending[g_npending].comm = comm; g_pending[g_npending].buf = recv; g_pending[g_npending].bytes = count * 4;
  g_pending[g_npending].root = root;
  ++g_npending; ++g_calls;
  return 0;
}
int ncclGroupEnd(void) {
  if (--g_depth > 0) return 0;
  /* every rank of a communicator must have registered exactly one broadcast with the same root and size */
  for (int i = 0; i < g_npending; ++i) {
    const PendingOp *src = NULL;
    int members = 0;
    for (int j = 0; j < g_npending; ++j)
      if (g_pending[j].comm->group_id == g_pending[i].comm->group_id) {
        ++members;
        if (g_pending[j].root != g_pending[i].root || g_pending[j].bytes != g_pending[i].bytes) { g_npending = 0; return 5; }
        if (g_pending[j].comm->rank == g_pending[j].root) src = &g_pending[j];
      }
    if (!src || members != g_pending[i].comm->nranks) { g_npending = 0; return 5; }
    if (g_pending[i].buf != src->buf) memcpy(g_pending[i].buf, src->buf, src->bytes);
  }
  g_npending = 0;
  return 0;
}
const char *ncclGetErrorString(int r) { return r == 5 ? "invalid usage (fake NCCL)" : "fake NCCL error"; }
