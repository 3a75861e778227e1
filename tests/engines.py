"""Test-side engines with the same surface as popscle_b200.capi.Context, so tests can drive the
product's host code (popscle_b200.cli) with the CPU oracle.  TEST INFRASTRUCTURE."""
import oracle_py as orc


class OracleEngine:
    def __init__(self, n_threads=8):
        self.n_threads = n_threads

    def demux_run(self, plp, gp, has_gp, alphas, doublet_prior=0.5, want_grid=False):
        return orc.demux(plp, gp, has_gp, alphas, doublet_prior, want_grid=want_grid, n_threads=self.n_threads)

    @staticmethod
    def fmx_opts(*a, **kw):
        return orc.fmx_opts(*a, **kw)

    def fmx_run(self, plp, opts, init_clust=None, want_clusters=False):
        r = orc.fmx_run(plp, opts, init_clust, want_clusters=want_clusters, n_threads=self.n_threads)
        return r["cells"], r["res"], r["clust_gl"], r["clust_cnt"]

    def fmx_run_aux(self, plp, opts, init_clust=None, compact=False):
        return orc.fmx_run_aux(plp, opts, init_clust, n_threads=self.n_threads)

    def close(self):
        pass
