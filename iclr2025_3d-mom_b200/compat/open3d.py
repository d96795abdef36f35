"""stub: only used by GaussianModel.grow() (opt.add_point, off by default)."""
