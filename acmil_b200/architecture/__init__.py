"""Same module paths as the reference's ``architecture`` package, backed by acmil_b200.heads."""
