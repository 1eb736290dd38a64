"""Built-in front end (csolve_b200/csrc/front.cpp): the flat model it produces equals the one
obtained from the reference's own structures after the reference's root phase (digests recorded
by tests/golden/make_golden.py through integration/csolve_gpu_shim.c)."""
import json
import os

import pytest

import csolve_b200 as cb
import util
from csolve_b200 import instances as I
from make_instances import instance_table, random_table

DIGESTS = json.load(open(os.path.join(util.GOLDEN, "flat_digests.json")))
ALL = dict(instance_table())
ALL.update(random_table())


def test_flat_models_equal_reference():
    n_ok = n_inf = 0
    for name, text in ALL.items():
        exp = DIGESTS[name]
        try:
            m = cb.Model(text)
        except cb.CsolveError as e:
            assert exp == {-3: "INFEASIBLE", -2: "SYNTAX"}.get(e.code), (name, e)
            n_inf += 1
            continue
        assert util.flat_digest(m.flat) == exp, name
        n_ok += 1
    assert n_ok >= 80 and n_inf >= 100


def test_live_reference_if_built():
    """when oracle/_ref exists (build container), compare flat models directly on fresh random inputs"""
    ref = util.reference_lib()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    import tempfile
    from gen_random import gen_instance
    same = 0
    for seed in range(900000, 900150):
        text = gen_instance(seed)
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(text)
        n = ref.ref_load(f.name.encode(), 0, 1)
        os.unlink(f.name)
        try:
            m = cb.Model(text)
        except cb.CsolveError as e:
            assert n < 0 and e.code in (-2, -3)
            continue
        assert n > 0 and ref.ref_flatten() == 0
        assert ref.ref_flat().contents.to_dict() == m.flat.to_dict()
        same += 1
    assert same > 20


def test_known_root_domains():
    """root domains the reference computes for schedule / wcet"""
    m = cb.Model(I.schedule())
    d = dict(zip(m.var_names, m.root_domains.reshape(-1, 2).tolist()))
    assert d["end"] == [11, 2147483646] and d["<obj>"] == [11, 2147483646]
    assert d["t1_start"] == [0, 1] and d["t2_start"] == [9, 15] and d["t3_end"] == [8, 9]
    m = cb.Model(I.wcet())
    d = dict(zip(m.var_names, m.root_domains.reshape(-1, 2).tolist()))
    assert d["<obj>"] == [-1697, 5392] and d["e4T"] == [0, 99] and d["e0"] == [1, 1] and d["m4F"] == [0, 200]
    assert m.var_names.index("<obj>") == 10          # objective expression's variables come first


def test_shapes_of_baseline_instances():
    """V / clauses with variables / nodes / W"""
    for text, shape in ((I.queens(8), (8, 84, 560, 168)), (I.queens(16), (16, 360, 2400, 720)),
                        (I.sudoku(I.SUDOKU_EXAMPLE), (81, 779, 3116, 1256)), (I.schedule(), (11, 7, 28, 14)),
                        (I.wcet(), (12, 15, 145, 50))):
        f = cb.Model(text).flat
        assert (f.n_vars, f.n_clauses, f.n_nodes, f.n_watch) == shape


def test_lexer_number_forms_and_comments():
    m = cb.Model("ALL;\n# comment\nx = 0b101 + 0x10 + 010 + 7; y = 0;\n")
    d = dict(zip(m.var_names, m.root_domains.reshape(-1, 2).tolist()))
    assert d["x"] == [5 + 16 + 8 + 7] * 2 and d["y"] == [0, 0]


@pytest.mark.parametrize("text,code", [
    ("ALL; x = ;", -2),                    # parser error
    ("ALL; x ? 3;", -2),                   # lexer error: invalid input
    ("", -2),                              # no objective
    ("ALL; 0 <= x; x <= 3; x > 5;", -3),   # INFEASIBLE PROBLEM at root (src/parser.y:71-73)
    ("ALL; 0 <= x;", -4),                  # unbounded variable: x (src/parser_support.c:249-251)
    ("ALL; 5;", -3),                       # propagating true into the constant 5 fails
])
def test_error_behaviour(text, code):
    with pytest.raises(cb.CsolveError) as e:
        cb.Model(text)
    assert e.value.code == code


def test_error_messages_follow_reference():
    for text, msg in (("ALL; x ? 3;", "invalid input `?' in line 1"), ("ALL;\n 0 <= x;", "unbounded variable: x")):
        with pytest.raises(cb.CsolveError) as e:
            cb.Model(text)
        assert msg in e.value.message


def test_weights_option():
    a = cb.Model(I.schedule(), compute_weights=True).flat.to_dict()["var_prio"]
    b = cb.Model(I.schedule(), compute_weights=False).flat.to_dict()["var_prio"]
    assert any(a) and not any(b)


def test_objective_override():
    assert cb.Model(I.queens(4, "ANY")).objective == cb.OBJ_ANY
    assert cb.Model(I.queens(4, "ANY"), objective=cb.OBJ_ALL).objective == cb.OBJ_ALL


def test_sudoku_batch_generator_is_deterministic_and_unique():
    a = I.sudoku_batch(3, seed=5)
    assert a == I.sudoku_batch(3, seed=5)
    for g in a:
        assert I._count_solutions(g) == 1
        r, _ = util.Oracle(cb.Model(I.sudoku(g))).solve_tree(cb.ORDER_SMALLEST_DOMAIN)
        assert r.solutions == 1
