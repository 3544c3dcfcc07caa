import sys
sys.path.insert(0,'/root/repo/qp-zk-circuits_b200')
import qpzk
ctx=qpzk.Context(0)
print('%.3e'%ctx.bench_poseidon(16))
