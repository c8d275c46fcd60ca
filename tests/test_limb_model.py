"""Limb-level model of the device field routines in circom_compat_b200/csrc/fp.cuh (CPU test, no GPU).

The CUDA code is straight-line PTX carry chains; this file restates the SAME sequences (mul4, cmad4, madc_shift,
madc_shift_m, chain1..4, the row order of mul_wide / redc / sqr_wide) with explicit 32-bit limbs and an explicit carry
flag, asserts that no closing add or top product ever overflows, and checks the results against Python integers.  It
pins the algorithm (index bookkeeping, operand-range preconditions); the GPU tests pin the transcription.

    mul_wide : eight row-shift rows, a < 2^255                      fp.cuh  Fp::prow / Fp::mul_wide
    redc     : eight reduction rows, m computed inside the chain    fp.cuh  Fp::madc_shift_m / Fp::mrow / Fp::redc
    sqr_wide : a^2 = sum_j a_j 2^(32j) [(a_j + msb(a_(j-1))) 2^(32j) + sum_(i<j) (2a)_i 2^(32i)], 36 products
                                                                     fp.cuh  Fp::chain1..4 / Fp::sqr_wide
"""
import random
M=0xffffffff
P=0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
INV=(-pow(P,-1,1<<32))%(1<<32)
def limbs(x,n=8): return [(x>>(32*i))&M for i in range(n)]
def val(l): return sum(v<<(32*i) for i,v in enumerate(l))
PL=limbs(P)

class CC:
    def __init__(s): s.c=0
cc=CC()
def add_cc(a,b):
    t=a+b; cc.c=t>>32; return t&M
def addc_cc(a,b):
    t=a+b+cc.c; cc.c=t>>32; return t&M
def addc(a,b):
    t=a+b+cc.c; assert t>>32==0, "overflow in closing addc"; return t&M
def mad_lo_cc(a,b,c):
    t=((a*b)&M)+c; cc.c=t>>32; return t&M
def madc_lo_cc(a,b,c):
    t=((a*b)&M)+c+cc.c; cc.c=t>>32; return t&M
def madc_hi_cc(a,b,c):
    t=((a*b)>>32)+c+cc.c; cc.c=t>>32; return t&M
def madc_hi(a,b,c):
    t=((a*b)>>32)+c+cc.c; assert t>>32==0,"overflow madc.hi"; return t&M

def mul4(x0,x1,x2,x3,b):
    out=[]
    for x in (x0,x1,x2,x3):
        p=x*b; out+= [p&M,p>>32]
    return out
def cmad4(acc,top,x0,x1,x2,x3,b):
    xs=(x0,x1,x2,x3)
    acc=list(acc)
    acc[0]=mad_lo_cc(xs[0],b,acc[0]); acc[1]=madc_hi_cc(xs[0],b,acc[1])
    for k in (1,2,3):
        acc[2*k]=madc_lo_cc(xs[k],b,acc[2*k]); acc[2*k+1]=madc_hi_cc(xs[k],b,acc[2*k+1])
    top=addc(top,0)
    return acc,top
def madc_shift(x0,e,a1,a3,a5,a7,b):
    e=list(e)
    x0=add_cc(x0,e[1])
    n=[0]*8
    n[0]=madc_lo_cc(a1,b,e[2]); n[1]=madc_hi_cc(a1,b,e[3])
    n[2]=madc_lo_cc(a3,b,e[4]); n[3]=madc_hi_cc(a3,b,e[5])
    n[4]=madc_lo_cc(a5,b,e[6]); n[5]=madc_hi_cc(a5,b,e[7])
    n[6]=madc_lo_cc(a7,b,0);    n[7]=madc_hi(a7,b,0)
    return x0,n
def madc_shift_m(x0,e,p1,p3,p5,p7):
    e=list(e)
    x0=add_cc(x0,e[1])
    m=(x0*INV)&M
    n=[0]*8
    n[0]=madc_lo_cc(p1,m,e[2]); n[1]=madc_hi_cc(p1,m,e[3])
    n[2]=madc_lo_cc(p3,m,e[4]); n[3]=madc_hi_cc(p3,m,e[5])
    n[4]=madc_lo_cc(p5,m,e[6]); n[5]=madc_hi_cc(p5,m,e[7])
    n[6]=madc_lo_cc(p7,m,0);    n[7]=madc_hi(p7,m,0)
    return x0,n,m

def mul_wide(a,b):
    a=limbs(a); b=limbs(b)
    t=[0]*16
    od=mul4(a[1],a[3],a[5],a[7],b[0]); ev=mul4(a[0],a[2],a[4],a[6],b[0]); t[0]=ev[0]
    x,e=od,ev   # next row: x = od (limb-0 aligned after retiring), e = ev
    for i in range(1,8):
        x0,e2=madc_shift(x[0],e,a[1],a[3],a[5],a[7],b[i]); x=[x0]+x[1:]
        x,top=cmad4(x,e2[7],a[0],a[2],a[4],a[6],b[i]); e2[7]=top
        t[i]=x[0]
        x,e=e2,x
    # now e = array whose [0] was just retired (limb-0 aligned, e[0] dropped), x = limb-1 aligned
    r=[0]*8
    r[0]=add_cc(x[0],e[1])
    for k in range(1,7): r[k]=addc_cc(x[k],e[k+1])
    r[7]=addc(x[7],0)
    t[8:]=r
    return t

def redc(t):
    x=list(t[:8]); 
    m=(x[0]*INV)&M
    e=mul4(PL[1],PL[3],PL[5],PL[7],m)
    x,top=cmad4(x,e[7],PL[0],PL[2],PL[4],PL[6],m); e[7]=top
    assert x[0]==0
    x,e=e,x
    for i in range(1,8):
        x0,e2,m=madc_shift_m(x[0],e,PL[1],PL[3],PL[5],PL[7]); x=[x0]+x[1:]
        x,top=cmad4(x,e2[7],PL[0],PL[2],PL[4],PL[6],m); e2[7]=top
        assert x[0]==0
        x,e=e2,x
    r=[0]*8
    r[0]=add_cc(x[0],e[1])
    for k in range(1,7): r[k]=addc_cc(x[k],e[k+1])
    r[7]=addc(x[7],0)
    r[0]=add_cc(r[0],t[8])
    for k in range(1,7): r[k]=addc_cc(r[k],t[8+k])
    r[7]=addc(r[7],t[15])
    v=val(r)
    return v-P if v>=P else v

def chain(acc,xs,b,add_last=0):
    # acc: list of 2*len(xs) limbs; first 2*(k-1) existing, last 2 fresh (ignored input)
    k=len(xs); acc=list(acc)
    if k==1:
        acc[0]=mad_lo_cc(xs[0],b,add_last); acc[1]=madc_hi(xs[0],b,0); return acc
    acc[0]=mad_lo_cc(xs[0],b,acc[0]); acc[1]=madc_hi_cc(xs[0],b,acc[1])
    for j in range(1,k-1):
        acc[2*j]=madc_lo_cc(xs[j],b,acc[2*j]); acc[2*j+1]=madc_hi_cc(xs[j],b,acc[2*j+1])
    acc[2*k-2]=madc_lo_cc(xs[k-1],b,add_last); acc[2*k-1]=madc_hi(xs[k-1],b,0)
    return acc

def sqr_wide(a):
    a=limbs(a)
    b=[(a[0]<<1)&M]+[((a[i]<<1)|(a[i-1]>>31))&M for i in range(1,7)]
    m=[0]+[a[j] if (a[j-1]>>31) else 0 for j in range(1,8)]
    ev=[None]*16; od=[None]*14
    # even class
    ev[0:2]=chain([0,0],[a[0]],a[0],0)
    ev[2:4]=chain([0,0],[a[1]],a[1],m[1])
    ev[2:6]=chain(ev[2:4]+[0,0],[b[0],a[2]],a[2],m[2])
    ev[4:8]=chain(ev[4:6]+[0,0],[b[1],a[3]],a[3],m[3])
    ev[4:10]=chain(ev[4:8]+[0,0],[b[0],b[2],a[4]],a[4],m[4])
    ev[6:12]=chain(ev[6:10]+[0,0],[b[1],b[3],a[5]],a[5],m[5])
    ev[6:14]=chain(ev[6:12]+[0,0],[b[0],b[2],b[4],a[6]],a[6],m[6])
    ev[8:16]=chain(ev[8:14]+[0,0],[b[1],b[3],b[5],a[7]],a[7],m[7])
    # odd class (od index = position-1)
    od[0:2]=chain([0,0],[b[0]],a[1])
    od[2:4]=chain([0,0],[b[1]],a[2])
    od[2:6]=chain(od[2:4]+[0,0],[b[0],b[2]],a[3])
    od[4:8]=chain(od[4:6]+[0,0],[b[1],b[3]],a[4])
    od[4:10]=chain(od[4:8]+[0,0],[b[0],b[2],b[4]],a[5])
    od[6:12]=chain(od[6:10]+[0,0],[b[1],b[3],b[5]],a[6])
    od[6:14]=chain(od[6:12]+[0,0],[b[0],b[2],b[4],b[6]],a[7])
    t=[0]*16
    t[0]=ev[0]
    t[1]=add_cc(ev[1],od[0])
    for k in range(2,15): t[k]=addc_cc(ev[k],od[k-1])
    t[15]=addc(ev[15],0)
    return t



def _rnd(rng, bound):
    if rng.random() < 0.3:      # limb patterns with extreme values
        return val([rng.choice([0, 1, M, M - 1, 0x80000000, 0x7fffffff, rng.getrandbits(32)]) for _ in range(8)]) % bound
    return rng.randrange(bound)


def test_mul_wide_rows():
    rng = random.Random(1)
    for _ in range(1500):
        a, b = _rnd(rng, 1 << 255), _rnd(rng, 1 << 256)
        assert val(mul_wide(a, b)) == a * b


def test_sqr_wide_identity():
    rng = random.Random(2)
    for a in [0, 1, (1 << 256) - 1, 1 << 255, (1 << 255) - 1] + [_rnd(rng, 1 << 256) for _ in range(1500)]:
        assert val(sqr_wide(a)) == a * a


def test_redc_rows():
    rng = random.Random(3)
    rinv = pow(1 << 256, -1, P)
    for _ in range(1500):
        x, y = _rnd(rng, P), _rnd(rng, P)
        assert redc(mul_wide(x, y)) == x * y * rinv % P
        t = _rnd(rng, P << 256)                     # anything below p * 2^256 (what mul_sub / the Fq2 routines feed it)
        assert redc(limbs(t, 16)) == t * rinv % P
