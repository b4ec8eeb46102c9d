import numpy as np
rng = np.random.default_rng(0)
def seq(x, vx, z, vz, th, om, tm, alpha, fx, fz, dt, steps):
    st, ct = np.sin(th), np.cos(th)
    adt, eps = alpha*dt, alpha*dt*dt
    d0 = om*dt
    sd, cd, se, ce = np.sin(d0), np.cos(d0), np.sin(eps), np.cos(eps)
    tmdt, fxdt, fzdt = tm*dt, fx*dt, fz*dt
    for s in range(steps):
        vx += st*tmdt + fxdt; vz += ct*tmdt + fzdt; om += adt
        x += vx*dt; z += vz*dt; th += om*dt
        sd, cd = sd*ce + cd*se, cd*ce - sd*se
        st, ct = st*cd + ct*sd, ct*cd - st*sd
    return x, vx, z, vz, th, om

def cmul(a, b):  # complex as (s, c) = (imag, real): rotation composition
    return (a[0]*b[1] + a[1]*b[0], a[1]*b[1] - a[0]*b[0])

def blocked(x, vx, z, vz, th, om, tm, alpha, fx, fz, dt, steps):
    # steps must be a multiple of 4
    st0, ct0 = np.sin(th), np.cos(th)
    adt, eps = alpha*dt, alpha*dt*dt
    d0 = om*dt
    zd = (np.sin(d0), np.cos(d0)); ce1 = (np.sin(eps), np.cos(eps))
    ce2 = cmul(ce1, ce1); ce3 = cmul(ce2, ce1); ce4 = cmul(ce2, ce2)
    # rotation increments of the 4 sub-steps of a block starting at delta state zd (= e^{i delta_k}): delta_{k+j} = delta_k + j eps
    w = (st0, ct0)      # e^{i theta_k} at block start
    S1s = S1c = S2s = S2c = 0.0
    n = steps
    k = 0
    while k < steps:
        z1 = cmul(zd, ce1); z2 = cmul(zd, ce2); z3 = cmul(zd, ce3); z4 = cmul(zd, ce4)   # e^{i delta_{k+1..k+4}}: depth 1 from zd
        u12 = cmul(z1, z2); u34 = cmul(z3, z4)          # depth 2
        u123 = cmul(u12, z3); u1234 = cmul(u12, u34)    # depth 3
        w1 = cmul(w, z1); w2 = cmul(w, u12); w3 = cmul(w, u123); w4 = cmul(w, u1234)      # theta_{k+1..k+4}: depth 1 from w (u's come from the zd chain)
        # sub-steps k..k+3 use sin/cos of theta_k, theta_{k+1}, theta_{k+2}, theta_{k+3} = w, w1, w2, w3
        for j, ww in enumerate((w, w1, w2, w3)):
            wt = n - (k + j)           # weight of a_{k+j} in the position sum
            S1s += ww[0]; S1c += ww[1]; S2s += wt*ww[0]; S2c += wt*ww[1]
        w = w4; zd = z4; k += 4
    tmdt, fxdt, fzdt = tm*dt, fx*dt, fz*dt
    T2 = n*(n+1)/2
    vxn = vx + tmdt*S1s + n*fxdt
    vzn = vz + tmdt*S1c + n*fzdt
    xn = x + dt*(n*vx + tmdt*S2s + fxdt*T2)
    zn = z + dt*(n*vz + tmdt*S2c + fzdt*T2)
    omn = om + n*adt
    thn = th + dt*(n*om + adt*T2)
    return xn, vxn, zn, vzn, thn, omn

worst = 0
for trial in range(2000):
    x, vx, z, vz = rng.uniform(-2,2,4); th = rng.uniform(-1.4,1.4); om = rng.uniform(-15,15)
    tm = rng.uniform(2, 22); alpha = rng.uniform(-400, 400); fx, fz = rng.uniform(-0.5,0.5), -9.8
    a = np.array(seq(x,vx,z,vz,th,om,tm,alpha,fx,fz,1e-3,20)); b = np.array(blocked(x,vx,z,vz,th,om,tm,alpha,fx,fz,1e-3,20))
    worst = max(worst, np.abs(a-b).max()/ (1+np.abs(a).max()))
print('worst rel diff', worst)


def sums(x, vx, z, vz, th, om, tm, alpha, fx, fz, dt, steps):
    '''the form the kernel uses: running sums S1 (of sin / cos theta_k) and P (of S1), closed-form velocities / positions / angle'''
    st, ct = np.sin(th), np.cos(th)
    adt, eps = alpha*dt, alpha*dt*dt
    d0 = om*dt
    sd, cd, se, ce = np.sin(d0), np.cos(d0), np.sin(eps), np.cos(eps)
    S1s = S1c = Ps = Pc = 0.0
    for s in range(steps):
        S1s += st; S1c += ct; Ps += S1s; Pc += S1c
        sd, cd = sd*ce + cd*se, cd*ce - sd*se
        st, ct = st*cd + ct*sd, ct*cd - st*sd
    n = steps; t2 = 0.5*n*(n+1)
    xn = x + dt*(n*vx + dt*(tm*Ps + fx*t2)); zn = z + dt*(n*vz + dt*(tm*Pc + fz*t2))
    return xn, vx + dt*(tm*S1s + n*fx), zn, vz + dt*(tm*S1c + n*fz), th + dt*(n*om + adt*t2), om + n*adt


worst = 0
for trial in range(2000):
    x, vx, z, vz = rng.uniform(-2,2,4); th = rng.uniform(-1.4,1.4); om = rng.uniform(-15,15)
    tm = rng.uniform(2, 22); alpha = rng.uniform(-400, 400); fx, fz = rng.uniform(-0.5,0.5), -9.8
    a = np.array(seq(x,vx,z,vz,th,om,tm,alpha,fx,fz,1e-3,20)); b = np.array(sums(x,vx,z,vz,th,om,tm,alpha,fx,fz,1e-3,20))
    worst = max(worst, np.abs(a-b).max()/ (1+np.abs(a).max()))
print('running-sum form: worst rel diff', worst)
