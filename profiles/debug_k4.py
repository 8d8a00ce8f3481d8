import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fpt_b200.api as api, fpt_b200.synth as synth
asize=bsize=500; regend,wsize,nsnp=200000,50000,600
ch=synth.chromosome(900+nsnp, regend, nsnp, asize, bsize)
res={}
for tag,mode in (("a2",2),("b2",2),("a1",1),("b1",1)):
    api.set_k4_mode(mode)
    res[tag]=api.css_scan(ch["acodes"],ch["bcodes"],ch["pos"],asize,bsize,regend,wsize,wsize,20,100,mds=0,seed=21,probes=True)
api.set_k4_mode(2)
for x,y in (("a2","b2"),("a1","b1"),("a2","a1")):
    X0,X1=res[x][3]["X"],res[y][3]["X"]
    print(x,y,"max|dX|",np.nanmax(np.abs(X0-X1)),"windows differing",[int((X0[w]!=X1[w]).sum()) for w in range(X0.shape[0])], "scores equal", np.array_equal(res[x][0],res[y][0]))
