"""Stand-in for the absent `emcee` package (test infrastructure, see ../README.md).

Restates the published affine-invariant ensemble sampler (Goodman & Weare 2010) with emcee 3.x
defaults: StretchMove(a=2), two splits with a shuffled half/half partition per step,
z = ((a-1)u+1)^2/a, proposal c - (c-s) z, accept iff (ndim-1) ln z + dlogp > ln u,
get_chain(flat, thin, discard) = chain[discard+thin-1::thin] reshaped step-major.
PARITY UNPINNED: the reference's tests pin only output sizes at this boundary.
"""
import numpy as np

LAST_SAMPLER = None  # make_golden.py reads the log-probability closure from here


class EnsembleSampler:

    def __init__(self, nwalkers, ndim, log_prob_fn, a=2.0, seed=None):
        global LAST_SAMPLER
        if nwalkers % 2:
            raise ValueError('nwalkers must be even')
        self.nwalkers, self.ndim, self.log_prob_fn, self.a = nwalkers, ndim, log_prob_fn, a
        self._random = np.random.RandomState(seed)
        self._chain = None
        self._log_prob = None
        LAST_SAMPLER = self

    def run_mcmc(self, initial_state, nsteps, progress=False, progress_kwargs=None):
        x = np.array(initial_state, dtype=float)
        nw, nd = x.shape
        lp = np.array([self.log_prob_fn(xi) for xi in x])
        chain = np.empty((nsteps, nw, nd))
        logp = np.empty((nsteps, nw))
        rng = self._random
        for t in range(nsteps):
            inds = np.arange(nw) % 2
            rng.shuffle(inds)
            for split in range(2):
                active = np.where(inds == split)[0]
                comp = x[inds != split]
                zz = ((self.a - 1.0) * rng.rand(active.size) + 1)**2.0 / self.a
                rint = rng.randint(comp.shape[0], size=active.size)
                q = comp[rint] - (comp[rint] - x[active]) * zz[:, None]
                new_lp = np.array([self.log_prob_fn(qi) for qi in q])
                lnpdiff = (nd - 1.0) * np.log(zz) + new_lp - lp[active]
                acc = lnpdiff > np.log(rng.rand(active.size))
                x[active[acc]] = q[acc]
                lp[active[acc]] = new_lp[acc]
            chain[t] = x
            logp[t] = lp
        self._chain, self._log_prob = chain, logp
        return x

    def get_chain(self, flat=False, thin=1, discard=0):
        v = self._chain[discard + thin - 1::thin]
        if flat:
            return v.reshape(-1, self.ndim)
        return v

    def get_log_prob(self, flat=False, thin=1, discard=0):
        v = self._log_prob[discard + thin - 1::thin]
        return v.reshape(-1) if flat else v
