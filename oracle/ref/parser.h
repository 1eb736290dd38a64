/* Stand-in for the bison-generated src/parser.h (bison is not available in the
 * build image). main.c only needs the yyparse() prototype from it. */
#ifndef STANDIN_PARSER_H
#define STANDIN_PARSER_H
int yyparse(void);
#endif
