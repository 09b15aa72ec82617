"""Import shim for scikit-sparse (absent offline).  Test infrastructure only."""
