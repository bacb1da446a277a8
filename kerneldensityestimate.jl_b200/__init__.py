"""B200-native hot path of KernelDensityEstimate.jl; import it as `kde_b200` (see ../kde_b200)."""
