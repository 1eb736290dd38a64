import random, sys, os
def gen_expr(rng, vars_, depth):
    if depth<=0 or rng.random()<0.3:
        if rng.random()<0.6: return rng.choice(vars_)
        return str(rng.randint(0,9))
    op=rng.choice(['+','-','*','neg','paren'])
    if op=='neg': 
        return '-'+ (rng.choice(vars_) if rng.random()<0.5 else str(rng.randint(0,5)))
    if op=='paren': return '('+gen_expr(rng,vars_,depth-1)+')'
    return gen_expr(rng,vars_,depth-1)+' '+op+' '+gen_expr(rng,vars_,depth-1)
def gen_rel(rng, vars_):
    r=rng.choice(['=','!=','<','<=','>','>='])
    return gen_expr(rng,vars_,2)+' '+r+' '+gen_expr(rng,vars_,2)
def gen_bool(rng, vars_, depth):
    if depth<=0 or rng.random()<0.4: 
        s=gen_rel(rng,vars_)
        if rng.random()<0.2: s='!('+s+')'
        return s
    op=rng.choice(['&','|'])
    return '('+gen_bool(rng,vars_,depth-1)+') '+op+' ('+gen_bool(rng,vars_,depth-1)+')'
def gen_instance(seed):
    rng=random.Random(seed+int(os.environ.get("SEED0","0")))
    nv=rng.randint(2,6)
    vars_=['v%d'%i for i in range(nv)]
    obj=rng.choice(['ANY;','ALL;','ALL;','MIN '+gen_expr(rng,vars_,2)+';','MAX '+gen_expr(rng,vars_,2)+';'])
    lines=[obj]
    for v in vars_:
        lo=rng.randint(-3,3); hi=lo+rng.randint(0,6)
        lines.append('%d <= %s; %s <= %d;'%(lo,v,v,hi))
    for _ in range(rng.randint(1,int(os.environ.get("NC","6")))):
        if rng.random()<0.15 and nv>=3:
            k=rng.randint(2,nv)
            args=[gen_expr(rng,vars_,1) for _ in range(k)]
            lines.append('all_different('+', '.join(args)+');')
        else:
            lines.append(gen_bool(rng,vars_,2)+';')
    rng.shuffle(lines[1:])
    return '\n'.join(lines)+'\n'
if __name__=='__main__':
    out=sys.argv[1]; n=int(sys.argv[2]); os.makedirs(out,exist_ok=True)
    for s in range(n):
        open(os.path.join(out,'r%05d.txt'%s),'w').write(gen_instance(s))
