import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/qp-zk-circuits_b200'); sys.path.insert(0,'/root/repo/tests')
from conftest import splitmix_columns
from oracle import oracle as orc
import qpzk
ctx=qpzk.Context(0)
def run(log_n, Cs, pow_bits, nq, fpb, salted=()):
    n=1<<log_n
    cols=[splitmix_columns(77+i+log_n,c,n) for i,c in enumerate(Cs)]
    rng=np.random.default_rng(5)
    salts=[rng.integers(0,orc.P,size=(4,n<<3),dtype=np.uint64) if i in salted else None for i in range(len(Cs))]
    gb=[]; ob=[]
    for i,c in enumerate(cols):
        if i==len(cols)-1:
            gb.append(qpzk.PolynomialBatch.from_coeffs(ctx,c,3,salts[i],4)); ob.append(orc.PolynomialBatch.from_coeffs(c,3,4,salts[i]))
        else:
            gb.append(qpzk.PolynomialBatch.from_values(ctx,c,3,salts[i],4)); ob.append(orc.PolynomialBatch.from_values(c,3,4,salts[i]))
    gch=qpzk.Challenger(ctx); och=orc.Challenger()
    digest=np.arange(4,dtype=np.uint64)+11
    gch.observe_elements(digest); och.observe(digest)
    for g,o in zip(gb,ob):
        assert (g.cap==o.cap()).all()
        gch.observe_cap(g); och.observe(o.cap())
    zg=gch.get_extension_challenge(); zo=och.get(2)
    assert (zg==zo).all(), (zg,zo)
    wn=orc.lib().orc_gl_root_of_unity(log_n)
    gz=[orc.lib().orc_gl_mul(int(zo[0]),wn), orc.lib().orc_gl_mul(int(zo[1]),wn)]
    allp=[(o,i) for o in range(len(Cs)) for i in range(Cs[o])]
    nxt=[(2,i) for i in range(min(2,Cs[2]))] if len(Cs)>2 else [(0,0)]
    batches=[(list(map(int,zo)),allp),(gz,nxt)]
    gi=qpzk.FriInstanceInfo(batches); oi=orc.Instance(batches)
    op=np.concatenate([np.concatenate([ob[o].eval_ext(zo) for o in range(len(Cs))]), np.stack([ob[o].eval_ext(gz)[i] for o,i in nxt])])
    gop=np.concatenate([gb[o].eval_ext(zg) for o in range(len(Cs))])
    assert (gop==op[:len(allp)]).all()
    gch.observe_elements(op); och.observe(op)
    gp=qpzk.fri_params(log_n,3,4,pow_bits,nq,4,fpb); opar=orc.fri_params(log_n,3,4,pow_bits,nq,4,fpb)
    assert gp.reduction_arity_bits==list(opar.arity_bits)[:opar.num_rounds]
    ochv=och.clone()
    t0=time.time(); gproof,gopen=qpzk.PolynomialBatch.prove_openings(gi,gb,gch,gp,return_openings=True); t1=time.time()
    oproof=orc.fri_prove(ob,oi,opar,och); t2=time.time()
    assert (gopen==op).all()
    same=(gproof==oproof).all()
    if not same:
        bad=np.nonzero(gproof!=oproof)[0]; print('  first mismatches',bad[:10], gproof.size)
    rc=orc.fri_verify(log_n,[b.leaf_len for b in gb],[b.cap for b in gb],oi,op,opar,ochv,gproof)
    st=(gch.export_state()==och.export())
    print('fri log_n',log_n,Cs,'rounds',gp.reduction_arity_bits,'same',same,'verify rc',rc,'chal state',st,'pow',gproof[-1],'gpu %.3fs cpu %.3fs'%(t1-t0,t2-t1))
run(6,[5,9,3,2],8,5,2)
run(5,[3,4,2,2],4,3,5)
run(8,[4,6,3,2],10,7,3,salted=(1,2,3))
run(10,[8,13,4,4],16,28,5)
run(13,[85,135,20,16],16,28,5)
