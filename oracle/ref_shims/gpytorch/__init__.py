"""Stand-in for gpytorch (absent in this image). ORACLE SCAFFOLDING ONLY: lets the
unmodified reference under /root/reference import inside the build container so its
outputs can pin oracle/cr_oracle.py and generate tests/golden/. Never on a product path."""
