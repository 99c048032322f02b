/*
 * TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for Intel's OpenMP runtime
 * (libiomp5.so) so that the reference's prebuilt ifort library
 * /root/reference/mlatom/fortran/np2/KREG.so can be dlopen()ed in this image, which has no
 * Intel runtime.  KREG.so imports exactly seven __kmpc_* entry points, all with symbol
 * version "VERSION" (see `nm -D KREG.so | grep kmpc`); they implement
 * `!$OMP PARALLEL DO ... SCHEDULE(STATIC)` (KREG.f90:312-344, 573-593).
 *
 * Threading model: __kmpc_fork_call spawns OMP_NUM_THREADS pthreads per parallel region
 * and joins them; the static-schedule init splits the iteration space into contiguous
 * blocks.  Nothing here is part of the shipped product.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>

typedef int32_t kmp_i32;
typedef void (*kmpc_micro)(kmp_i32 *gtid, kmp_i32 *btid, ...);

#define MAX_SHARED 32

static __thread kmp_i32 t_id = 0;   /* this thread's index in its team */
static __thread kmp_i32 t_team = 1; /* team size */

static int team_size_from_env(void) {
  const char *s = getenv("OMP_NUM_THREADS");
  int n = s ? atoi(s) : 1;
  if (n < 1) n = 1;
  if (n > 256) n = 256;
  return n;
}

kmp_i32 __kmpc_global_thread_num(void *loc) { (void)loc; return t_id; }
kmp_i32 __kmpc_ok_to_fork(void *loc) { (void)loc; return 1; }
void __kmpc_serialized_parallel(void *loc, kmp_i32 gtid) { (void)loc; (void)gtid; }
void __kmpc_end_serialized_parallel(void *loc, kmp_i32 gtid) { (void)loc; (void)gtid; }
void __kmpc_for_static_fini(void *loc, kmp_i32 gtid) { (void)loc; (void)gtid; }

/* schedule(static) without chunk: contiguous blocks, remainder spread over the first threads */
void __kmpc_for_static_init_4(void *loc, kmp_i32 gtid, kmp_i32 sched, kmp_i32 *plast,
                              kmp_i32 *plower, kmp_i32 *pupper, kmp_i32 *pstride,
                              kmp_i32 incr, kmp_i32 chunk) {
  (void)loc; (void)gtid; (void)sched; (void)incr; (void)chunk;
  long lo = *plower, hi = *pupper;
  long n = hi - lo + 1;
  if (n < 0) n = 0;
  long base = n / t_team, extra = n % t_team;
  long mine = base + (t_id < extra ? 1 : 0);
  long start = lo + (long)t_id * base + (t_id < extra ? t_id : extra);
  *plower = (kmp_i32)start;
  *pupper = (kmp_i32)(start + mine - 1);
  if (plast) *plast = (t_id == t_team - 1);
  if (pstride) *pstride = (kmp_i32)n;
}

struct team_job {
  kmpc_micro fn;
  void *shared[MAX_SHARED];
  kmp_i32 id, team;
};

static void *team_thread(void *p) {
  struct team_job *j = (struct team_job *)p;
  kmp_i32 saved_id = t_id, saved_team = t_team;
  t_id = j->id;
  t_team = j->team;
  kmp_i32 g = j->id, b = j->id;
  void **a = j->shared;
  /* surplus pointer arguments are harmless under the SysV x86-64 calling convention */
  j->fn(&g, &b, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12],
        a[13], a[14], a[15], a[16], a[17], a[18], a[19], a[20], a[21], a[22], a[23], a[24], a[25],
        a[26], a[27], a[28], a[29], a[30], a[31]);
  t_id = saved_id;
  t_team = saved_team;
  return NULL;
}

void __kmpc_fork_call(void *loc, kmp_i32 argc, kmpc_micro fn, ...) {
  (void)loc;
  int T = team_size_from_env();
  if (argc > MAX_SHARED) abort();
  void *shared[MAX_SHARED] = {0};
  va_list ap;
  va_start(ap, fn);
  for (int i = 0; i < argc; i++) shared[i] = va_arg(ap, void *);
  va_end(ap);

  struct team_job *jobs = (struct team_job *)calloc((size_t)T, sizeof *jobs);
  pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof *th);
  for (int t = 0; t < T; t++) {
    jobs[t].fn = fn;
    jobs[t].id = t;
    jobs[t].team = T;
    for (int i = 0; i < MAX_SHARED; i++) jobs[t].shared[i] = shared[i];
  }
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 64u << 20); /* ifort puts automatic arrays on the stack */
  for (int t = 1; t < T; t++) pthread_create(&th[t], &attr, team_thread, &jobs[t]);
  team_thread(&jobs[0]);
  for (int t = 1; t < T; t++) pthread_join(th[t], NULL);
  pthread_attr_destroy(&attr);
  free(jobs);
  free(th);
}
