"""Oracle package: CPU restatements of the reference path.  TEST INFRASTRUCTURE ONLY (see reference_port.py)."""
