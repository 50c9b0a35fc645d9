import sys, ctypes, numpy as np
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import fftlibs as fl
PROD=fl.Lib(fl.product()); ORC=fl.Lib(fl.oracle(),"orc_")
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 7)
lengths=[2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,19,20,21,22,23,24,25,26,27,28,29,30,31,32,33,34,35,37,39,40,41,43,45,47,48,49,50,51,53,55,59,60,61,63,64,65,66,67,70,71,72,73,75,77,79,80,81,83,84,85,87,88,89,90,91,96,97,99,100,101,104,105,108,110,111,112,113,117,119,120,121,125,126,127,128,129,130,131,132,135,137,139,140,143,144,147,149,150,151,153,154,156,157,160,162,163,165,167,168,169,170,171,173,175,176,179,180,181,182,187,189,191,192,193,195,196,197,198,199,200,208,210,216,220,221,224,225,231,234,240,242,243,245,250,252,255,256,257,260,264,270,272,273,275,280,286,288,289,294,297,300,308,312,315,320,323,324,325,330,336,338,340,343,350,351,352,357,360,361,363,364,374,375,378,384,385,390,392,396,400,405,416,420,425,429,432,440,441,442,448,450,455,459,462,468,476,480,486,490,495,500,504,507,510,512,1000,1001,1002,1023,1024,1025,1536,2000,2048,2187,2310,2401,2500,3003,3072,3125,4000,4095,4096,4097,4913,5005,6000,6561,6859,7777,8000,8191,8192,8193,9999,10000,12288,15625,16384,17017,19683,20000,30030,32768,50000,65536]
worst=0; nbad=0
for case in range(400):
    fam=fl.FAMILIES[int(rng.integers(len(fl.FAMILIES)))]
    n=int(lengths[int(rng.integers(len(lengths)))])
    lot=int(rng.choice([1,2,3,4,5,7,8,9,15,16,17,31,32,33,63,64,65,100,127,128,129]))
    if n*lot>600000: lot=max(1,600000//n)
    layout=int(rng.integers(5))
    if layout==0: inc,jump=1,n
    elif layout==1: inc,jump=1,n+int(rng.integers(1,9))
    elif layout==2: inc,jump=lot,1
    elif layout==3: inc,jump=lot+int(rng.integers(1,4)),1
    else: inc=int(rng.integers(2,4)); jump=n*inc+int(rng.integers(0,5))
    d="fb"[int(rng.integers(2))]
    cnt=(lot-1)*jump+(n-1)*inc+1
    x=fl.rand_input(fam,cnt,case)
    a,ia=PROD.runm(fam,d,lot,jump,n,inc,x,work=False)
    b,ib=ORC.runm(fam,d,lot,jump,n,inc,x)
    if ia!=ib: print("IER",fam,n,lot,inc,jump,d,ia,ib); nbad+=1; continue
    if ia: continue
    e=fl.rel_l2(a,b); lim=fl.tol(n)+fl.ref_noise(fam,n)*0+ (3e-14 if fl.max_generic_factor(fl.underlying(fam,n))>13 else 0)
    worst=max(worst,e)
    if e>lim: print("BAD",fam,n,lot,inc,jump,d,e,lim); nbad+=1
print("cases 400 worst",worst,"bad",nbad)
