"""Instance tables shared by the golden generator and the tests (seeded, deterministic)."""
from csolve_b200 import instances as I
from gen_random import gen_instance


def instance_table():
    return {
        "queens4": I.queens(4), "queens6": I.queens(6), "queens8": I.queens(8), "queens8any": I.queens(8, "ANY"),
        "queens10": I.queens(10), "sudoku": I.sudoku(I.SUDOKU_EXAMPLE), "sudoku_any": I.sudoku(I.SUDOKU_EXAMPLE, "ANY"),
        "schedule": I.schedule(), "wcet": I.wcet(),
        "sat20all": I.random_3sat(20, seed=1, objective="ALL"), "sat50": I.random_3sat(50, seed=1),
        "sat50all": I.random_3sat(50, seed=2, objective="ALL"), "sat100": I.random_3sat(100, seed=1),
    }


def random_table(n=300, base=500000):
    return {"rand%04d" % i: gen_instance(base + i) for i in range(n)}
