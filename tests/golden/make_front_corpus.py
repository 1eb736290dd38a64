#!/usr/bin/env python3
"""Front-end hardening corpus: the reference's own fuzz seeds (fuzz/inputs/*.txt) and examples,
mutated with the tokens of its fuzz dictionary (fuzz/dict) the way afl-fuzz's dictionary stage does (token insert /
overwrite, span delete / duplicate, byte flips), each mutant run through the COMPILED REFERENCE CLI (oracle/_ref/csolve_ref,
-c false -r 0 so that ALL counts are exact) and recorded with what the reference did:

    kind     "ok" | "syntax" (lexer / parser error) | "infeasible" (INFEASIBLE PROBLEM) | "unbounded" | "fatal"
    message  the reference's own error text after "error: " (exact for lexer / unbounded / invalid-operation errors;
             parser messages come from bison, which the image lacks, so only their kind and line number are comparable)
    result   for "ok": SOLUTIONS, last BEST, "NO SOLUTION FOUND"

Run in the build container only:  python tests/golden/make_front_corpus.py   -> tests/golden/front_corpus.json
"""
import json
import os
import random
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

REF = "/root/reference"
N_MUTANTS = 2000


def dictionary():
    toks = []
    for line in open(os.path.join(REF, "fuzz", "dict")):
        m = re.match(r'\w+="(.*)"', line.strip())
        if m:
            toks.append(m.group(1))
    return toks + ["0", "1", "7", "-1", "2147483647", "-2147483648", "99999999999", "0x1f", "0b101", "017", "x1", "y", "\n", " ", "#"]


def mutate(rng, text, toks):
    b = list(text)
    for _ in range(1 if rng.random() < 0.6 else rng.randint(2, 4)):
        op = rng.choice([0, 0, 1, 1, 2, 3, 3, 4, 5])
        pos = rng.randrange(len(b) + 1)
        if op == 0:
            b[pos:pos] = list(rng.choice(toks))                         # dictionary insert
        elif op == 1 and b:
            t = list(rng.choice(toks)); b[pos:pos + len(t)] = t          # dictionary overwrite
        elif op == 2 and b:
            q = min(len(b), pos + rng.randint(1, 12)); del b[pos:q]      # delete a span
        elif op == 3 and b:
            q = min(len(b), pos + rng.randint(1, 24)); b[pos:pos] = b[pos:q]   # duplicate a span
        elif op == 4 and b:
            b[min(pos, len(b) - 1)] = chr(rng.randint(33, 126))          # byte overwrite
        else:
            b = b[:pos] if pos > 8 else b                                # truncate
    return "".join(b)


def run_reference(text):
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(text)
        path = f.name
    try:
        p = subprocess.run([util.REF_CLI, "-s", "0", "-c", "false", "-r", "0", "-t", "5", path], capture_output=True, text=True, timeout=20,
                           errors="replace")
    except subprocess.TimeoutExpired:
        os.unlink(path)
        return None
    os.unlink(path)
    out, err = p.stdout, p.stderr
    m = re.search(r"error: (.*)", err)
    if m:
        msg = m.group(1).strip()
        if msg.startswith("invalid input"):
            kind = "syntax"
        elif msg.startswith("unbounded variable"):
            kind = "unbounded"
        elif re.search(r" in line \d+$", msg):
            kind = "syntax"
        else:
            kind = "fatal"
        return dict(kind=kind, message=msg)
    if "INFEASIBLE PROBLEM" in out:
        return dict(kind="infeasible")
    if "TIMEOUT" in out:
        return None
    st = re.search(r"SOLUTIONS: (\d+)", out)
    if st is None:
        return dict(kind="fatal", message="rc=%d" % p.returncode)
    best = re.findall(r"BEST: (-?\d+)", out)
    return dict(kind="ok", solutions=int(st.group(1)), best=int(best[-1]) if best else None, no_solution="NO SOLUTION FOUND" in out)


def main():
    assert os.path.exists(util.REF_CLI), "build oracle/_ref first (make -C oracle/ref)"
    toks = dictionary()
    seeds = []
    for d, names in (("fuzz/inputs", sorted(os.listdir(os.path.join(REF, "fuzz", "inputs")))), ("examples", ["queens8.txt", "schedule.txt"])):
        for n in names:
            seeds.append(open(os.path.join(REF, d, n)).read())
    rng = random.Random(20261018)
    corpus, seen = [], set()
    for s in seeds:                                   # the seeds themselves first
        r = run_reference(s)
        if r is not None:
            corpus.append(dict(text=s, **r)); seen.add(s)
    while len(corpus) < N_MUTANTS:
        t = mutate(rng, rng.choice(seeds), toks)
        if t in seen or "\x00" in t:
            continue
        seen.add(t)
        r = run_reference(t)
        if r is None:
            continue                                   # the reference itself timed out: not a front-end case
        corpus.append(dict(text=t, **r))
    kinds = {}
    for c in corpus:
        kinds[c["kind"]] = kinds.get(c["kind"], 0) + 1
    json.dump(corpus, open(os.path.join(HERE, "front_corpus.json"), "w"), indent=0)
    print(len(corpus), kinds)


if __name__ == "__main__":
    main()
