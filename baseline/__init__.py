"""Reference arm of bench.py: the UNMODIFIED reference model code, copied into the git-ignored baseline/_ref/ by
baseline/make_ref.py and imported by baseline/ref_loader.py. Nothing under nsdp_b200/ imports this package."""
