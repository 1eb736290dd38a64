/*
 * csolve_grammar.h -- hand-written lexer + recursive-descent parser for the
 * csolve input language.
 *
 * flex and bison are not available in the build image, so the reference's
 * generated front end (src/lexer.l, src/parser.y) cannot be produced. This
 * header is the stand-in: it recognises the same tokens (lexer.l:36-102) and
 * the same grammar (parser.y:94-284) and performs the same sequence of
 * semantic actions, but through a small builder interface so that it can
 * drive either
 *   - the product front end (csolve_b200/csrc/front.cpp, own expression tree), or
 *   - the reference engine's own constr_t/env_t structures (oracle/ref/standin_parser.c).
 *
 * The includer defines, BEFORE including this file:
 *
 *   CSG_EXPR                      expression handle type (pointer-like; CSG_NULL is "none")
 *   CSG_NULL
 *   CSG_EXPR csg_num(void *ctx, int32_t v);                  PrimaryExpr: NUM      (parser.y:134-137)
 *   CSG_EXPR csg_ident(void *ctx, const char *name);         PrimaryExpr: IDENT    (parser.y:138-147)
 *   CSG_EXPR csg_expr(void *ctx, int op, CSG_EXPR l, CSG_EXPR r);   new expression node; op is one of CSG_OP_*
 *   void     csg_weighten(void *ctx, CSG_EXPR e, int weight_class);  (parser.y:219-265) weight_class = CSG_W_*
 *   CSG_EXPR csg_wand_new(void *ctx);                        all_different: empty wide-and (parser.y:169-171)
 *   void     csg_wand_append(void *ctx, CSG_EXPR w, CSG_EXPR e);     (parser.y:179-182)
 *   CSG_EXPR csg_objective(void *ctx, int kind, CSG_EXPR e); Objective rule (parser.y:109-131); returns clause 0
 *   void     csg_constraint(void *ctx, CSG_EXPR e);          Constraints rule (parser.y:94-99)
 *   void     csg_error(void *ctx, int is_lexer, int ch, const char *msg, unsigned line);  must not return normally
 *                                                            for lexer errors (lexer.l:98-101 exits)
 *
 * Entry point: int csg_parse(void *ctx, const char *text, size_t len)
 *   returns 0 on success, 1 after a syntax error (as yyparse does).
 */
#ifndef CSOLVE_GRAMMAR_H
#define CSOLVE_GRAMMAR_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* expression operators as the grammar sees them (same characters as csolve.h:133-155) */
#define CSG_OP_EQ  '='
#define CSG_OP_LT  '<'
#define CSG_OP_NEG '-'
#define CSG_OP_ADD '+'
#define CSG_OP_MUL '*'
#define CSG_OP_NOT '!'
#define CSG_OP_AND '&'
#define CSG_OP_OR  '|'

/* weight classes (parser_support.h:22-27) */
#define CSG_W_EQUAL     1000
#define CSG_W_COMPARE   100
#define CSG_W_NOT_EQUAL 10

/* objective kinds in the order of enum objective_t (csolve.h:241-247) */
#define CSG_OBJ_ANY 0
#define CSG_OBJ_ALL 1
#define CSG_OBJ_MIN 2
#define CSG_OBJ_MAX 3

enum csg_tok {
  CSG_T_EOF = 0, CSG_T_ANY, CSG_T_ALL, CSG_T_MIN, CSG_T_MAX, CSG_T_ALLDIFF,
  CSG_T_NEQ, CSG_T_LEQ, CSG_T_GEQ, CSG_T_NUM, CSG_T_IDENT,
  CSG_T_CHAR /* single-character token, value in .ch */
};

struct csg_state {
  void *ctx;
  const char *p, *end;
  unsigned line;
  enum csg_tok tok;
  int ch;            /* for CSG_T_CHAR */
  int32_t num;       /* for CSG_T_NUM */
  char *ident;       /* for CSG_T_IDENT (NUL-terminated, reused buffer) */
  size_t ident_cap;
  int failed;
};

static int csg_is_sym_start(int c) {
  return c == '_' || c == '@' || c == '$' || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
}
static int csg_is_sym(int c) { return csg_is_sym_start(c) || (c >= '0' && c <= '9'); }
static int csg_is_xdigit(int c) {
  return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F');
}

/* convert like the lexer actions do: strtol, then narrowing to domain_t (lexer.l:70-88) */
static int32_t csg_strtol(const char *s, size_t n, int base) {
  char buf[80];
  if (n >= sizeof(buf)) {
    /* very long literal: strtol saturates to LONG_MAX, narrowing gives -1 */
    char *big = (char *)malloc(n + 1);
    memcpy(big, s, n); big[n] = 0;
    long v = strtol(big, NULL, base);
    free(big);
    return (int32_t)v;
  }
  memcpy(buf, s, n); buf[n] = 0;
  return (int32_t)strtol(buf, NULL, base);
}

static void csg_next(struct csg_state *s) {
  for (;;) {
    if (s->p >= s->end) { s->tok = CSG_T_EOF; return; }
    int c = (unsigned char)*s->p;
    if (c == '\n') { s->line++; s->p++; continue; }
    if (c == ' ' || c == '\t' || c == '\r') { s->p++; continue; }
    if (c == '#') { while (s->p < s->end && *s->p != '\n') s->p++; continue; }
    break;
  }
  const char *b = s->p;
  int c = (unsigned char)*b;
  size_t left = (size_t)(s->end - b);

  if (csg_is_sym_start(c)) {
    const char *e = b + 1;
    while (e < s->end && csg_is_sym((unsigned char)*e)) e++;
    size_t n = (size_t)(e - b);
    s->p = e;
    if (n == 3 && memcmp(b, "ANY", 3) == 0) { s->tok = CSG_T_ANY; return; }
    if (n == 3 && memcmp(b, "ALL", 3) == 0) { s->tok = CSG_T_ALL; return; }
    if (n == 3 && memcmp(b, "MAX", 3) == 0) { s->tok = CSG_T_MAX; return; }
    if (n == 3 && memcmp(b, "MIN", 3) == 0) { s->tok = CSG_T_MIN; return; }
    if (n == 13 && memcmp(b, "all_different", 13) == 0) { s->tok = CSG_T_ALLDIFF; return; }
    if (n + 1 > s->ident_cap) {
      s->ident_cap = 2 * (n + 1);
      s->ident = (char *)realloc(s->ident, s->ident_cap);
    }
    memcpy(s->ident, b, n); s->ident[n] = 0;
    s->tok = CSG_T_IDENT;
    return;
  }
  if (c >= '0' && c <= '9') {
    /* longest match among BNUM / ONUM / DNUM / XNUM (lexer.l:36-39) */
    if (c == '0') {
      if (left >= 3 && b[1] == 'b' && (b[2] == '0' || b[2] == '1')) {
        const char *e = b + 2;
        while (e < s->end && (*e == '0' || *e == '1')) e++;
        s->num = csg_strtol(b + 2, (size_t)(e - b - 2), 2);
        s->p = e; s->tok = CSG_T_NUM; return;
      }
      if (left >= 3 && (b[1] == 'x' || b[1] == 'X') && csg_is_xdigit((unsigned char)b[2])) {
        const char *e = b + 2;
        while (e < s->end && csg_is_xdigit((unsigned char)*e)) e++;
        s->num = csg_strtol(b + 2, (size_t)(e - b - 2), 16);
        s->p = e; s->tok = CSG_T_NUM; return;
      }
      const char *e = b + 1;
      while (e < s->end && *e >= '0' && *e <= '7') e++;
      s->num = csg_strtol(b, (size_t)(e - b), 8);
      s->p = e; s->tok = CSG_T_NUM; return;
    }
    const char *e = b + 1;
    while (e < s->end && *e >= '0' && *e <= '9') e++;
    s->num = csg_strtol(b, (size_t)(e - b), 10);
    s->p = e; s->tok = CSG_T_NUM; return;
  }
  if (c == '!' && left >= 2 && b[1] == '=') { s->p += 2; s->tok = CSG_T_NEQ; return; }
  if (c == '<' && left >= 2 && b[1] == '=') { s->p += 2; s->tok = CSG_T_LEQ; return; }
  if (c == '>' && left >= 2 && b[1] == '=') { s->p += 2; s->tok = CSG_T_GEQ; return; }
  switch (c) {
  case '=': case '<': case '>': case '-': case '+': case '*': case '!': case '&': case '|':
  case '(': case ')': case ',': case ';':
    s->p++; s->tok = CSG_T_CHAR; s->ch = c; return;
  default:
    csg_error(s->ctx, 1, c, NULL, s->line);
    /* a lexer error is fatal in the reference (lexer.l:98-101); if the
       builder returns anyway, stop parsing */
    s->failed = 1; s->tok = CSG_T_EOF; s->p = s->end;
    return;
  }
}

static int csg_is_char(struct csg_state *s, int ch) { return s->tok == CSG_T_CHAR && s->ch == ch; }

static void csg_syntax_error(struct csg_state *s, const char *msg) {
  if (!s->failed) {
    s->failed = 1;
    csg_error(s->ctx, 0, 0, msg, s->line);
  }
}

static CSG_EXPR csg_parse_expr(struct csg_state *s);

/* PrimaryExpr (parser.y:133-151) */
static CSG_EXPR csg_parse_primary(struct csg_state *s) {
  if (s->failed) return CSG_NULL;
  if (s->tok == CSG_T_NUM) {
    CSG_EXPR e = csg_num(s->ctx, s->num);
    csg_next(s);
    return e;
  }
  if (s->tok == CSG_T_IDENT) {
    CSG_EXPR e = csg_ident(s->ctx, s->ident);
    csg_next(s);
    return e;
  }
  if (csg_is_char(s, '(')) {
    csg_next(s);
    CSG_EXPR e = csg_parse_expr(s);
    if (s->failed) return CSG_NULL;
    if (!csg_is_char(s, ')')) { csg_syntax_error(s, "syntax error, expecting ')'"); return CSG_NULL; }
    csg_next(s);
    return e;
  }
  csg_syntax_error(s, "syntax error, expecting NUM or IDENT or '('");
  return CSG_NULL;
}

/* UnaryExpr (parser.y:153-186): '-' and '!' apply to a PrimaryExpr only */
static CSG_EXPR csg_parse_unary(struct csg_state *s) {
  if (s->failed) return CSG_NULL;
  if (csg_is_char(s, '-') || csg_is_char(s, '!')) {
    int op = s->ch == '-' ? CSG_OP_NEG : CSG_OP_NOT;
    csg_next(s);
    CSG_EXPR p = csg_parse_primary(s);
    if (s->failed) return CSG_NULL;
    return csg_expr(s->ctx, op, p, CSG_NULL);
  }
  if (s->tok == CSG_T_ALLDIFF) {
    csg_next(s);
    if (!csg_is_char(s, '(')) { csg_syntax_error(s, "syntax error, expecting '('"); return CSG_NULL; }
    csg_next(s);
    /* ExprList builds a list by prepending (parser_support.c:275-284), so the
       pair loops of parser.y:173-183 walk the arguments last-to-first */
    size_t n = 0, cap = 16;
    CSG_EXPR *list = (CSG_EXPR *)malloc(cap * sizeof(CSG_EXPR));
    for (;;) {
      CSG_EXPR e = csg_parse_expr(s);
      if (s->failed) { free(list); return CSG_NULL; }
      if (n == cap) { cap *= 2; list = (CSG_EXPR *)realloc(list, cap * sizeof(CSG_EXPR)); }
      list[n++] = e;
      if (csg_is_char(s, ',')) { csg_next(s); continue; }
      break;
    }
    if (!csg_is_char(s, ')')) { free(list); csg_syntax_error(s, "syntax error, expecting ',' or ')'"); return CSG_NULL; }
    csg_next(s);
    CSG_EXPR w = csg_wand_new(s->ctx);
    for (size_t li = n; li-- > 0; ) {        /* l walks the reversed list */
      for (size_t ki = li; ki-- > 0; ) {     /* k = l->next ... */
        CSG_EXPR a = csg_expr(s->ctx, CSG_OP_EQ, list[li], list[ki]);
        CSG_EXPR b = csg_expr(s->ctx, CSG_OP_NOT, a, CSG_NULL);
        csg_wand_append(s->ctx, w, b);
      }
    }
    free(list);
    return w;
  }
  return csg_parse_primary(s);
}

/* MultExpr (parser.y:197-202) */
static CSG_EXPR csg_parse_mult(struct csg_state *s) {
  CSG_EXPR l = csg_parse_unary(s);
  while (!s->failed && csg_is_char(s, '*')) {
    csg_next(s);
    CSG_EXPR r = csg_parse_unary(s);
    if (s->failed) return CSG_NULL;
    l = csg_expr(s->ctx, CSG_OP_MUL, l, r);
  }
  return l;
}

/* AddExpr (parser.y:204-215): a-b is ADD(a, NEG(b)) */
static CSG_EXPR csg_parse_add(struct csg_state *s) {
  CSG_EXPR l = csg_parse_mult(s);
  while (!s->failed && (csg_is_char(s, '+') || csg_is_char(s, '-'))) {
    int minus = s->ch == '-';
    csg_next(s);
    CSG_EXPR r = csg_parse_mult(s);
    if (s->failed) return CSG_NULL;
    if (minus) r = csg_expr(s->ctx, CSG_OP_NEG, r, CSG_NULL);
    l = csg_expr(s->ctx, CSG_OP_ADD, l, r);
  }
  return l;
}

/* RelatExpr (parser.y:217-249) */
static CSG_EXPR csg_parse_relat(struct csg_state *s) {
  CSG_EXPR l = csg_parse_add(s);
  for (;;) {
    if (s->failed) return CSG_NULL;
    int kind;
    if (csg_is_char(s, '<')) kind = 0;
    else if (csg_is_char(s, '>')) kind = 1;
    else if (s->tok == CSG_T_LEQ) kind = 2;
    else if (s->tok == CSG_T_GEQ) kind = 3;
    else break;
    csg_next(s);
    CSG_EXPR r = csg_parse_add(s);
    if (s->failed) return CSG_NULL;
    CSG_EXPR e;
    switch (kind) {
    case 0: e = csg_expr(s->ctx, CSG_OP_LT, l, r); break;                 /* l < r */
    case 1: e = csg_expr(s->ctx, CSG_OP_LT, r, l); break;                 /* l > r == r < l */
    case 2: e = csg_expr(s->ctx, CSG_OP_NOT, csg_expr(s->ctx, CSG_OP_LT, r, l), CSG_NULL); break; /* !(r < l) */
    default: e = csg_expr(s->ctx, CSG_OP_NOT, csg_expr(s->ctx, CSG_OP_LT, l, r), CSG_NULL); break; /* !(l < r) */
    }
    csg_weighten(s->ctx, e, CSG_W_COMPARE);
    l = e;
  }
  return l;
}

/* EqualExpr (parser.y:251-268) */
static CSG_EXPR csg_parse_equal(struct csg_state *s) {
  CSG_EXPR l = csg_parse_relat(s);
  for (;;) {
    if (s->failed) return CSG_NULL;
    int neq;
    if (csg_is_char(s, '=')) neq = 0;
    else if (s->tok == CSG_T_NEQ) neq = 1;
    else break;
    csg_next(s);
    CSG_EXPR r = csg_parse_relat(s);
    if (s->failed) return CSG_NULL;
    CSG_EXPR e = csg_expr(s->ctx, CSG_OP_EQ, l, r);
    if (neq) {
      e = csg_expr(s->ctx, CSG_OP_NOT, e, CSG_NULL);
      csg_weighten(s->ctx, e, CSG_W_NOT_EQUAL);
    } else {
      csg_weighten(s->ctx, e, CSG_W_EQUAL);
    }
    l = e;
  }
  return l;
}

/* AndExpr (parser.y:270-275) */
static CSG_EXPR csg_parse_and(struct csg_state *s) {
  CSG_EXPR l = csg_parse_equal(s);
  while (!s->failed && csg_is_char(s, '&')) {
    csg_next(s);
    CSG_EXPR r = csg_parse_equal(s);
    if (s->failed) return CSG_NULL;
    l = csg_expr(s->ctx, CSG_OP_AND, l, r);
  }
  return l;
}

/* OrExpr / Expr (parser.y:277-285) */
static CSG_EXPR csg_parse_expr(struct csg_state *s) {
  CSG_EXPR l = csg_parse_and(s);
  while (!s->failed && csg_is_char(s, '|')) {
    csg_next(s);
    CSG_EXPR r = csg_parse_and(s);
    if (s->failed) return CSG_NULL;
    l = csg_expr(s->ctx, CSG_OP_OR, l, r);
  }
  return l;
}

static int csg_expect_semi(struct csg_state *s) {
  if (s->failed) return 0;
  if (!csg_is_char(s, ';')) { csg_syntax_error(s, "syntax error, expecting ';'"); return 0; }
  csg_next(s);
  return 1;
}

/* Input := Objective Constraint*  (parser.y:53-131) */
static int csg_parse(void *ctx, const char *text, size_t len) {
  struct csg_state st;
  memset(&st, 0, sizeof(st));
  st.ctx = ctx; st.p = text; st.end = text + len; st.line = 1;
  struct csg_state *s = &st;
  csg_next(s);
  if (!s->failed) {
    switch (s->tok) {
    case CSG_T_ANY: csg_next(s); if (csg_expect_semi(s)) csg_objective(ctx, CSG_OBJ_ANY, CSG_NULL); break;
    case CSG_T_ALL: csg_next(s); if (csg_expect_semi(s)) csg_objective(ctx, CSG_OBJ_ALL, CSG_NULL); break;
    case CSG_T_MIN: case CSG_T_MAX: {
      int kind = s->tok == CSG_T_MIN ? CSG_OBJ_MIN : CSG_OBJ_MAX;
      csg_next(s);
      CSG_EXPR e = csg_parse_expr(s);
      if (csg_expect_semi(s)) csg_objective(ctx, kind, e);
      break;
    }
    default:
      csg_syntax_error(s, "syntax error, expecting ANY or ALL or MIN or MAX");
    }
  }
  while (!s->failed && s->tok != CSG_T_EOF) {
    CSG_EXPR e = csg_parse_expr(s);
    if (csg_expect_semi(s)) csg_constraint(ctx, e);
  }
  free(st.ident);
  return st.failed ? 1 : 0;
}

#endif /* CSOLVE_GRAMMAR_H */
