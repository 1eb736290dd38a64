import sys, random, os
sys.path.insert(0,os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I
seeds=[I.queens(5), I.sudoku(I.SUDOKU_EXAMPLE), I.schedule(), I.wcet(), I.random_3sat(12, seed=3), "MIN 3*a - 2*b + c;\n0 <= a; a <= 10; 0 <= b; b <= 10; 0 <= c; c <= 5; a + b <= 12; a != b;\n",
       "ALL;\nall_different(a, b, c);\n1 <= a; a <= 3; 1 <= b; b <= 3; 1 <= c; c <= 3;\n", "ANY;\nx = 0x1f; y = 0b101; z = 017; x < y | !(y < z) & z != 3;\n"]
toks=["(",")",";","+","-","*","<","<=",">",">=","=","!=","!","&","|",",","all_different","MIN","MAX","ANY","ALL","0","1","-1","2147483647","-2147483648","99999999999","x","y","#",
      "\n"," ","0x","0b","a1"]
rng=random.Random(int(sys.argv[1]))
ok=err=0
for it in range(int(sys.argv[2])):
    s=rng.choice(seeds)
    b=list(s)
    for _ in range(rng.randint(1,6)):
        op=rng.randint(0,4)
        pos=rng.randrange(len(b)+1)
        if op==0 and b: del b[min(pos,len(b)-1)]
        elif op==1: b[pos:pos]=list(rng.choice(toks))
        elif op==2 and b: b[min(pos,len(b)-1)]=chr(rng.randint(1,126))
        elif op==3: b=b[:pos]
        else:
            q=rng.randrange(len(b)+1); a,c=min(pos,q),max(pos,q); b[a:c]=b[a:c][::-1]
    text="".join(b)
    try:
        m=cb.Model(text); ok+=1; m.close()
    except cb.CsolveError as e:
        err+=1
print("seed",sys.argv[1],"parsed",ok,"rejected",err)
