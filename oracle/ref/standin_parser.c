/*
 * standin_parser.c -- stand-in for the reference's flex/bison front end, used
 * ONLY to build the reference oracle (oracle/_ref). TEST INFRASTRUCTURE.
 *
 * Compiled against the reference's own headers (-I /root/reference/src) and
 * linked with the reference's unmodified engine objects. It provides the
 * symbols the generated lexer.c/parser.c would provide (yyparse, yyset_in,
 * yyget_in, yylex_destroy) and performs the semantic actions of
 * src/parser.y:53-285 through the shared grammar skeleton
 * csolve_b200/csrc/csolve_grammar.h.
 */
#include "csolve.h"
#include "parser_support.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CSG_EXPR struct constr_t *
#define CSG_NULL NULL

static struct constr_t *_root = NULL;   /* the Constraints wide-and */

static struct constr_t *csg_num(void *ctx, int32_t v) {
  (void)ctx;
  struct constr_t *c = alloc(sizeof(struct constr_t));
  *c = CONSTRAINT_TERM(VALUE(v));
  return c;
}

static struct constr_t *csg_ident(void *ctx, const char *name) {
  (void)ctx;
  struct env_t *var = vars_find_key(name);
  if (var != NULL) {
    return var->val;
  }
  struct constr_t *c = alloc(sizeof(struct constr_t));
  *c = CONSTRAINT_TERM(INTERVAL(DOMAIN_MIN, DOMAIN_MAX));
  vars_add(name, c);
  return c;
}

static struct constr_t *csg_expr(void *ctx, int op, struct constr_t *l, struct constr_t *r) {
  (void)ctx;
  struct constr_t *c = alloc(sizeof(struct constr_t));
  switch (op) {
  case '=': *c = CONSTRAINT_EXPR(EQ, l, r); break;
  case '<': *c = CONSTRAINT_EXPR(LT, l, r); break;
  case '-': *c = CONSTRAINT_EXPR(NEG, l, NULL); break;
  case '+': *c = CONSTRAINT_EXPR(ADD, l, r); break;
  case '*': *c = CONSTRAINT_EXPR(MUL, l, r); break;
  case '!': *c = CONSTRAINT_EXPR(NOT, l, NULL); break;
  case '&': *c = CONSTRAINT_EXPR(AND, l, r); break;
  case '|': *c = CONSTRAINT_EXPR(OR, l, r); break;
  default: print_fatal(ERROR_MSG_INVALID_OPERATION, op);
  }
  return c;
}

static void csg_weighten(void *ctx, struct constr_t *e, int weight_class) {
  (void)ctx;
  if (strategy_compute_weights()) {
    vars_weighten(e, weight_class / max(1, vars_count(e)));
  }
}

static struct constr_t *csg_wand_new(void *ctx) {
  (void)ctx;
  struct constr_t *c = alloc(sizeof(struct constr_t));
  *c = CONSTRAINT_WAND(0, NULL);
  return c;
}

static void csg_wand_append(void *ctx, struct constr_t *w, struct constr_t *e) {
  (void)ctx;
  w->constr.wand.length++;
  const size_t size = w->constr.wand.length * sizeof(struct wand_expr_t);
  w->constr.wand.elems = realloc(w->constr.wand.elems, size);
  w->constr.wand.elems[w->constr.wand.length-1] =
    (struct wand_expr_t) { .constr = e, .orig = e, .prop_tag = 0 };
}

static struct constr_t *csg_objective(void *ctx, int kind, struct constr_t *e) {
  (void)ctx;
  struct constr_t *c = alloc(sizeof(struct constr_t));
  switch (kind) {
  case 0:
    objective_init(OBJ_ANY, &shared()->objective_best);
    *c = CONSTRAINT_TERM(VALUE(1));
    break;
  case 1:
    objective_init(OBJ_ALL, &shared()->objective_best);
    *c = CONSTRAINT_TERM(VALUE(1));
    break;
  case 2:
    objective_init(OBJ_MIN, &shared()->objective_best);
    vars_add("<obj>", objective_val());
    *c = CONSTRAINT_EXPR(EQ, e, objective_val());
    break;
  default:
    objective_init(OBJ_MAX, &shared()->objective_best);
    vars_add("<obj>", objective_val());
    *c = CONSTRAINT_EXPR(EQ, objective_val(), e);
    break;
  }
  _root = alloc(sizeof(struct constr_t));
  *_root = CONSTRAINT_WAND(0, NULL);
  csg_wand_append(NULL, _root, c);
  return c;
}

static void csg_constraint(void *ctx, struct constr_t *e) {
  csg_wand_append(ctx, _root, e);
}

static void csg_error(void *ctx, int is_lexer, int ch, const char *msg, unsigned line) {
  (void)ctx;
  if (is_lexer) {
    print_error(ERROR_MSG_LEXER_ERROR, ch, line);
    exit(EXIT_FAILURE);
  }
  print_error(ERROR_MSG_PARSER_ERROR, msg, line);
}

#include "csolve_grammar.h"

static FILE *_in = NULL;

void yyset_in(FILE *f) { _in = f; }
FILE *yyget_in(void) { return _in; }
int yylex_destroy(void) { return 0; }

static char *slurp(FILE *f, size_t *len) {
  size_t cap = 1 << 16, n = 0;
  char *buf = malloc(cap);
  for (;;) {
    size_t r = fread(buf + n, 1, cap - n, f);
    n += r;
    if (r == 0) break;
    if (n == cap) { cap *= 2; buf = realloc(buf, cap); }
  }
  *len = n;
  return buf;
}

/* result of the root phase, for the replay/flatten library */
size_t standin_size = 0;
struct env_t *standin_env = NULL;
struct constr_t *standin_norm = NULL;
struct constr_t *standin_root = NULL;
/* when non-zero, yyparse() stops after the root phase and keeps everything alive */
int standin_stop_after_root = 0;
/* 0: ok, 1: syntax error, 2: infeasible at root */
int standin_status = 0;

/* The Input action of parser.y:55-92 */
int yyparse(void) {
  size_t len;
  char *text = slurp(_in != NULL ? _in : stdin, &len);
  _root = NULL;
  int rc = csg_parse(NULL, text, len);
  free(text);
  standin_status = rc;
  if (rc != 0 || _root == NULL) {
    return 1;
  }

  size_t size = var_count();

  prop_result_t prop = propagate(_root, size);
  struct constr_t *norm = _root;

  if (prop != PROP_ERROR) {
    struct constr_t *prev;
    do {
      prev = norm;
      norm = normalize(norm);
      prop = propagate(norm, size);
    } while (norm != prev && prop != PROP_ERROR);
  }

  if (prop == PROP_ERROR) {
    fprintf(stdout, "INFEASIBLE PROBLEM\n");
    standin_status = 2;
  }

  bind_commit();
  patch_commit();

  stats_init();

  if (prop != PROP_ERROR) {
    struct env_t *env = env_generate();

    clauses_init(norm, NULL);
    strategy_var_order_init(size, env);

    standin_size = size;
    standin_env = env;
    standin_norm = norm;
    standin_root = _root;
    if (standin_stop_after_root) {
      return 0;
    }

    solve(size, env, norm);

    env_free();
  }

  expr_free(_root);
  return 0;
}
