"""Instance generators for the BASELINE.json configurations.

All generators return csolve input text (grammar: reference src/parser.y,
tokens: src/lexer.l). Nothing here reads /root/reference: the GPU box does not
have it.

* ``queens``       same text as reference scripts/gen_queens.sh writes for N
* ``sudoku``       27 all_different groups + clues, the structure of examples/sudoku.txt
* ``sudoku_batch`` seeded generator of unique-solution puzzles (config 2; the
                   reference has no generator)
* ``schedule``     the 3-task schedule model of examples/schedule.txt (MIN)
* ``wcet``         the WCET/branch-prediction ILP of examples/wcet.txt (MAX)
* ``random_3sat``  uniform random 3-SAT, written the way scripts/cnf2csolve writes a DIMACS file
"""
import random

# the puzzle of reference examples/sudoku.txt, row-major
SUDOKU_EXAMPLE = "..53.....8......2..7..1.5..4....53...1..7...6..32...8..6.5....9..4....3......97.."

# a solved grid used as the seed of sudoku_batch()
_SOLVED = (
    "534678912" "672195348" "198342567"
    "859761423" "426853791" "713924856"
    "961537284" "287419635" "345286179"
)


def queens(n, objective="ALL"):
    """N-queens as scripts/gen_queens.sh emits it (the script's header is ANY)."""
    xs = ["X%d" % i for i in range(1, n + 1)]
    lines = ["# N-queens problem for N=%d" % n, "%s;" % objective]
    lines.append("all_different(" + ", ".join(xs) + ");")
    lines.append("all_different(" + ", ".join("X%d+%d" % (i, i) for i in range(1, n + 1)) + ");")
    lines.append("all_different(" + ", ".join("X%d-%d" % (i, i) for i in range(1, n + 1)) + ");")
    for i in range(1, n + 1):
        lines.append("1 <= X%d; X%d <= %d;" % (i, i, n))
    return "\n".join(lines) + "\n"


def _cell(r, c):
    """cell name: box letter A..I, position 1..9 inside the box (as examples/sudoku.txt)"""
    return "%s%d" % ("ABCDEFGHI"[(r // 3) * 3 + c // 3], (r % 3) * 3 + c % 3 + 1)


def sudoku(grid, objective="ALL"):
    """grid: 81 characters row-major, '.' or '0' for empty."""
    assert len(grid) == 81
    lines = ["%s;" % objective, "", "# Initial values"]
    # clues in box order, as the example lists them
    clues = []
    for r in range(9):
        for c in range(9):
            ch = grid[r * 9 + c]
            if ch not in ".0":
                clues.append((_cell(r, c), ch))
    for name, ch in sorted(clues):
        lines.append("%s = %s;" % (name, ch))
    lines += ["", "# Rows"]
    for r in range(9):
        lines.append("all_different(" + ", ".join(_cell(r, c) for c in range(9)) + ");")
    lines += ["", "# Columns"]
    for c in range(9):
        lines.append("all_different(" + ", ".join(_cell(r, c) for r in range(9)) + ");")
    lines += ["", "# Boxes"]
    for b in range(9):
        lines.append("all_different(" + ", ".join("%s%d" % ("ABCDEFGHI"[b], p) for p in range(1, 10)) + ");")
    lines += ["", "# Value ranges"]
    for b in range(9):
        for p in range(1, 10):
            n = "%s%d" % ("ABCDEFGHI"[b], p)
            lines.append("1 <= %s; %s <= 9;" % (n, n))
    return "\n".join(lines) + "\n"


def sudoku_roots(model_var_names, grids):
    """Root domain vectors [len(grids), 2 * 81] for the clue-free model sudoku('.' * 81): clue cells are
    single values, the others [1, 9]."""
    import numpy as np
    index = {n: k for k, n in enumerate(model_var_names)}
    out = np.empty((len(grids), 2 * len(model_var_names)), np.int32)
    out[:, 0::2] = 1
    out[:, 1::2] = 9
    for g, grid in enumerate(grids):
        for r in range(9):
            for c in range(9):
                ch = grid[r * 9 + c]
                if ch not in ".0":
                    k = index[_cell(r, c)]
                    out[g, 2 * k] = out[g, 2 * k + 1] = int(ch)
    return out


def _count_solutions(grid, limit=2):
    """tiny bitmask backtracker used only to keep generated puzzles unique (generator, not solver path)"""
    rows, cols, boxes = [0] * 9, [0] * 9, [0] * 9
    cells = []
    for i, ch in enumerate(grid):
        r, c = divmod(i, 9)
        if ch in ".0":
            cells.append(i)
        else:
            b = 1 << (int(ch) - 1)
            rows[r] |= b; cols[c] |= b; boxes[(r // 3) * 3 + c // 3] |= b
    count = 0

    def rec(todo):
        nonlocal count
        if count >= limit:
            return
        if not todo:
            count += 1
            return
        # most constrained cell first
        best_i, best_m, best_n = -1, 0, 10
        for k, i in enumerate(todo):
            r, c = divmod(i, 9)
            m = ~(rows[r] | cols[c] | boxes[(r // 3) * 3 + c // 3]) & 0x1FF
            n = bin(m).count("1")
            if n < best_n:
                best_i, best_m, best_n = k, m, n
                if n <= 1:
                    break
        if best_n == 0:
            return
        i = todo[best_i]
        rest = todo[:best_i] + todo[best_i + 1:]
        r, c = divmod(i, 9)
        bx = (r // 3) * 3 + c // 3
        m = best_m
        while m:
            b = m & -m
            m ^= b
            rows[r] |= b; cols[c] |= b; boxes[bx] |= b
            rec(rest)
            rows[r] ^= b; cols[c] ^= b; boxes[bx] ^= b

    rec(cells)
    return count


def sudoku_puzzle(rng, min_clues=26):
    """A unique-solution puzzle: permute the seed grid, then dig holes while uniqueness holds."""
    digits = list("123456789")
    rng.shuffle(digits)
    g = [[digits[int(_SOLVED[r * 9 + c]) - 1] for c in range(9)] for r in range(9)]
    # band / row-in-band / stack / column-in-stack permutations keep validity
    def perm_lines():
        bands = [0, 1, 2]
        rng.shuffle(bands)
        out = []
        for b in bands:
            rows = [0, 1, 2]
            rng.shuffle(rows)
            out += [b * 3 + r for r in rows]
        return out
    rp, cp = perm_lines(), perm_lines()
    g = [[g[rp[r]][cp[c]] for c in range(9)] for r in range(9)]
    if rng.random() < 0.5:
        g = [list(x) for x in zip(*g)]
    flat = [g[r][c] for r in range(9) for c in range(9)]
    order = list(range(81))
    rng.shuffle(order)
    clues = 81
    for i in order:
        if clues <= min_clues:
            break
        keep = flat[i]
        flat[i] = "."
        if _count_solutions("".join(flat)) != 1:
            flat[i] = keep
        else:
            clues -= 1
    return "".join(flat)


def _transform(grid, rng):
    """validity- and uniqueness-preserving symmetry: digit relabelling, band/stack/row/column permutations, transpose"""
    digits = list("123456789")
    rng.shuffle(digits)

    def lines():
        bands = [0, 1, 2]
        rng.shuffle(bands)
        out = []
        for b in bands:
            rows = [0, 1, 2]
            rng.shuffle(rows)
            out += [b * 3 + r for r in rows]
        return out
    rp, cp = lines(), lines()
    g = [[grid[rp[r] * 9 + cp[c]] for c in range(9)] for r in range(9)]
    if rng.random() < 0.5:
        g = [list(x) for x in zip(*g)]
    return "".join(ch if ch == "." else digits[int(ch) - 1] for row in g for ch in row)


def sudoku_batch(count, seed=20261018, min_clues=26, base=0):
    """`count` unique-solution puzzles as 81-character strings (deterministic in seed).
    base == 0: every puzzle is dug independently (slow: a uniqueness check per removed clue).
    base > 0 : `base` puzzles are dug, the rest are random symmetry transforms of them (distinct
               instances, uniqueness preserved) -- how the 10 000-instance batch of config 2 is made."""
    rng = random.Random(seed)
    if base <= 0 or base >= count:
        return [sudoku_puzzle(rng, min_clues) for _ in range(count)]
    dug = [sudoku_puzzle(rng, min_clues) for _ in range(base)]
    out = list(dug)
    while len(out) < count:
        out.append(_transform(dug[len(out) % base], rng))
    return out


def schedule():
    """3 tasks, release/WCET/deadline, precedences t1->t2, t1->t3, no overlap; minimise the makespan.
    The model of reference examples/schedule.txt (optimum 11)."""
    tasks = [("t1", 0, 3, 16), ("t2", 1, 2, 16), ("t3", 2, 4, 7)]
    lines = ["MIN end;"]
    for name, release, wcet, deadline in tasks:
        lines += ["%s_release = %d;" % (name, release),
                  "%s_release <= %s_start;" % (name, name),
                  "%s_end = %s_start + %d;" % (name, name, wcet),
                  "%s_end <= %s_release + %d;" % (name, name, deadline)]
    lines += ["t1_end <= t2_start;", "t1_end <= t3_start;"]
    for a, b in (("t1", "t2"), ("t1", "t3"), ("t2", "t3")):
        lines.append("%s_start > %s_end | %s_start > %s_end;" % (a, b, b, a))
    for name, _, _, _ in tasks:
        lines.append("end >= %s_end;" % name)
    return "\n".join(lines) + "\n"


def wcet():
    """Worst-case execution time with branch-prediction misses as an integer program (MAX).
    The model of reference examples/wcet.txt (optimum 1560)."""
    obj = " + ".join(["4*e1T", "-4*m1T", "6*m1T", "4*e1F", "-4*m1F", "6*m1F", "8*e2", "2*e3",
                      "3*e4T", "-3*m4T", "5*m4T", "3*e4F", "-3*m4F", "5*m4F"])
    lines = ["MAX %s;" % obj, "e0 = 1;", "e1T = e2;", "e1F = e3;", "e4T <= 99;",
             "e0 + e4T = e1T + e1F;", "e2 + e3 = e4T + e4F;",
             "m1T <= e1T;", "m1F <= e1F;", "m4T <= e4T;", "m4F <= e4F;",
             "m1T <= 14 * e0 + e1F + e4F;", "m1F <= 14 * e0 + e1T + e4T;",
             "m4T <= 20 * e0 + e1F + e4F;", "m1F <= 20 * e0 + e1T + e4T;",
             "m1T + m4T <= 28 * e0 + e1F + e4F;", "m1F + m4F <= 28 * e0 + e1T + e4T;"]
    for v in ["e0", "e1T", "e1F", "e2", "e3", "e4T", "e4F", "m1T", "m1F", "m4T", "m4F"]:
        lines.append("0 <= %s;" % v)
    return "\n".join(lines) + "\n"


def random_3sat_cnf(n, ratio=4.26, seed=1):
    """clauses of 3 distinct variables, each sign with p = 1/2"""
    rng = random.Random(seed)
    m = int(round(ratio * n))
    clauses = []
    for _ in range(m):
        vs = rng.sample(range(1, n + 1), 3)
        clauses.append([v if rng.random() < 0.5 else -v for v in vs])
    return clauses


def cnf_to_csolve(n, clauses, objective="ANY"):
    """what `awk -f scripts/cnf2csolve` prints for a DIMACS file: `a | !b | c;` lines, then 0/1 bounds"""
    lines = ["%s;" % objective]
    for cl in clauses:
        lines.append(" | ".join(("!x%d" % -l) if l < 0 else ("x%d" % l) for l in cl) + ";")
    for i in range(1, n + 1):
        lines.append("0 <= x%d; x%d <= 1;" % (i, i))
    return "\n".join(lines) + "\n"


def random_3sat(n, ratio=4.26, seed=1, objective="ANY"):
    return cnf_to_csolve(n, random_3sat_cnf(n, ratio, seed), objective)
