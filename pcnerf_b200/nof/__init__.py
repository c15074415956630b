"""Mirror of the reference's `nof` package for the ray-rendering hot path: same module paths below this package
(`nof.render`, `nof.networks`, `nof.criteria`, `nof.dataset.ipb2dmapping`), same names and signatures."""
