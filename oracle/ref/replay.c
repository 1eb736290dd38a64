/*
 * replay.c -- replay / inspection API over the UNMODIFIED reference engine.
 * TEST INFRASTRUCTURE (built into oracle/_ref/libcsolve_ref.so); never linked
 * into the product.
 *
 * ref_load()   : the option defaults of src/main.c:51-130 + the root phase of
 *                src/parser.y:55-85 (through the stand-in parser), stopping
 *                right before solve().
 * ref_replay() : one node transition exactly as solve() performs it:
 *                bind_level_set + step_enter (src/csolve.c:448-449,294-304),
 *                objective_update_val (src/objective.c:101-126),
 *                check_assignment (src/csolve.c:247-253); then undone with
 *                step_leave (src/csolve.c:307-314).
 * ref_flatten(): integration/csolve_gpu_shim.c applied to the reference's
 *                structures, for comparison with the built-in front end.
 */
#include "csolve.h"
#include "parser_support.h"
#include "csolve_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern size_t standin_size;
extern struct env_t *standin_env;
extern struct constr_t *standin_norm;
extern struct constr_t *standin_root;
extern int standin_stop_after_root;
extern int standin_status;
void yyset_in(FILE *);
int yyparse(void);

int csolve_flatten_reference(size_t size, struct env_t *env, struct constr_t *constr,
                             csolve_flat_model *out);
void csolve_flat_model_release(csolve_flat_model *m);

static int _loaded = 0;
static const char *_name = "csolve_ref";
const char *main_name(void) { return _name; }

static void ref_unload(void) {
  if (_loaded) {
    env_free();
    if (standin_root != NULL) expr_free(standin_root);
    /* the variable order only exists when the root phase succeeded */
    if (standin_env != NULL) strategy_var_order_free();
    bind_free(); patch_free(); alloc_free(); conflict_alloc_free();
    _loaded = 0;
    standin_env = NULL; standin_norm = NULL; standin_root = NULL; standin_size = 0;
  }
}

/* returns number of variables, -1 syntax error / cannot open, -2 infeasible at root */
int ref_load(const char *path, int create_conflicts, int compute_weights) {
  ref_unload();
  FILE *f = fopen(path, "r");
  if (f == NULL) return -1;
  bind_init(BIND_STACK_SIZE_DEFAULT);
  strategy_create_conflicts_init(create_conflicts != 0);
  strategy_prefer_failing_init(STRATEGY_PREFER_FAILING_DEFAULT);
  if (shared() == NULL) shared_init(WORKERS_MAX_DEFAULT);
  alloc_init(ALLOC_STACK_SIZE_DEFAULT);
  conflict_alloc_init(CONFLICT_ALLOC_STACK_SIZE_DEFAULT);
  strategy_order_init(STRATEGY_ORDER_DEFAULT);
  patch_init(PATCH_STACK_SIZE_DEFAULT);
  strategy_restart_frequency_init(STRATEGY_RESTART_FREQUENCY_DEFAULT);
  stats_frequency_init(0);
  timeout_init(TIME_MAX_DEFAULT);
  strategy_compute_weights_init(compute_weights != 0);
  shared()->solutions = 0;
  yyset_in(f);
  standin_stop_after_root = 1;
  standin_env = NULL;
  int rc = yyparse();
  fclose(f);
  _loaded = 1;
  if (rc != 0) { return -1; }
  if (standin_status == 2 || standin_env == NULL) { return -2; }
  return (int)standin_size;
}

int ref_nvars(void) { return (int)standin_size; }
const char *ref_var_name(int i) { return standin_env[i].key; }
int ref_objective(void) { return (int)objective(); }
int ref_obj_var(void) {
  if (objective_val() != NULL && objective_val()->constr.term.env != NULL)
    return (int)(objective_val()->constr.term.env - standin_env);
  return -1;
}

void ref_get_domains(int32_t *out) {
  for (size_t i = 0; i < standin_size; i++) {
    out[2*i] = standin_env[i].val->constr.term.val.lo;
    out[2*i+1] = standin_env[i].val->constr.term.val.hi;
  }
}

/* One node transition from an arbitrary state. dom_in == NULL keeps the current
 * (root) domains. Returns 1 if the node failed, 0 otherwise; dom_out receives the
 * post-propagation domains (meaningful only when not failed). The engine state is
 * restored afterwards. props_out (optional) receives the PROPS delta. */
int ref_replay(const int32_t *dom_in, int var, int32_t val, int32_t best,
               int32_t *dom_out, uint64_t *props_out) {
  size_t n = standin_size;
  struct val_t *saved = malloc(n * sizeof(struct val_t));
  for (size_t i = 0; i < n; i++) {
    saved[i] = standin_env[i].val->constr.term.val;
    if (dom_in != NULL) {
      standin_env[i].val->constr.term.val = INTERVAL(dom_in[2*i], dom_in[2*i+1]);
    }
  }
  domain_t saved_best = shared()->objective_best;
  if (objective() == OBJ_MIN || objective() == OBJ_MAX) {
    shared()->objective_best = best;
  }
  uint64_t props0 = stat_get_props();

  struct env_t *v = &standin_env[var];
  void *marker = alloc(0);
  size_t pdepth = patch(NULL, NULL);
  size_t bdepth = bind_depth();
  bind_level_set(0);
  if (!is_const(v->val)) {
    bind(v, VALUE(val), NULL);
  }
  /* objective_update_val() writes <obj> directly (not trailed): remember it */
  struct val_t obj_saved = objective_val()->constr.term.val;
  objective_update_val();

  int failed =
    propagate_clauses(&v->clauses) == PROP_ERROR ||
    (objective_val() != NULL && objective_val()->constr.term.env != NULL &&
     propagate_clauses(&objective_val()->constr.term.env->clauses) == PROP_ERROR);

  if (dom_out != NULL) ref_get_domains(dom_out);
  if (props_out != NULL) *props_out = stat_get_props() - props0;

  unbind(bdepth);
  unpatch(pdepth);
  dealloc(marker);
  objective_val()->constr.term.val = obj_saved;
  for (size_t i = 0; i < n; i++) {
    standin_env[i].val->constr.term.val = saved[i];
  }
  shared()->objective_best = saved_best;
  free(saved);
  return failed;
}

/* The reference's own solve() (src/csolve.c:398) on the loaded model, timed without process start, parsing or
 * printing: stdout is pointed at /dev/null for the duration (solve() prints every solution and the final stats).
 * Returns the seconds spent inside solve(); calls / solutions receive CALLS and SOLUTIONS. The engine state is
 * left as solve() leaves it: call ref_load() again before anything else. */
#include <fcntl.h>
#include <time.h>
#include <unistd.h>
double ref_solve_timed(uint64_t *calls, uint64_t *solutions) {
  fflush(stdout);
  int saved = dup(1);
  int devnull = open("/dev/null", O_WRONLY);
  dup2(devnull, 1);
  close(devnull);
  stats_init();
  shared()->solutions = 0;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  solve(standin_size, standin_env, standin_norm);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  fflush(stdout);
  dup2(saved, 1);
  close(saved);
  if (calls != NULL) *calls = stat_get_calls();
  if (solutions != NULL) *solutions = shared()->solutions;
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* leaf test of update_solution(): is_true(eval(root)) (src/csolve.c:226) */
int ref_eval_root(const int32_t *dom_in) {
  size_t n = standin_size;
  struct val_t *saved = malloc(n * sizeof(struct val_t));
  for (size_t i = 0; i < n; i++) {
    saved[i] = standin_env[i].val->constr.term.val;
    standin_env[i].val->constr.term.val = INTERVAL(dom_in[2*i], dom_in[2*i+1]);
  }
  int t = is_true(standin_norm->type->eval(standin_norm));
  for (size_t i = 0; i < n; i++) standin_env[i].val->constr.term.val = saved[i];
  free(saved);
  return t;
}

/* ---- flat model of the reference's structures, serialised as text ---------- */
static csolve_flat_model _flat;
static int _flat_valid = 0;

int ref_flatten(void) {
  if (_flat_valid) { csolve_flat_model_release(&_flat); _flat_valid = 0; }
  int rc = csolve_flatten_reference(standin_size, standin_env, standin_norm, &_flat);
  _flat_valid = rc == 0;
  return rc;
}
const csolve_flat_model *ref_flat(void) { return _flat_valid ? &_flat : NULL; }

/* ---- direct access to the reference's scalar functions (golden vectors) ---- */
int32_t ref_neg(int32_t a) { return neg(a); }
int32_t ref_add(int32_t a, int32_t b) { return add(a, b); }
int32_t ref_mul(int32_t a, int32_t b) { return mul(a, b); }
