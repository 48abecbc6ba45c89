"""Stand-in for scikit-image (absent from this image): only what jni/skan's csr.py imports.  See ../README.md."""
