"""Host-side glue shared by the two drop-in packages (``rdesign``, ``rnampnn``): Lightning callbacks / trainer
factory and small file readers.  Nothing here is on the B200 hot path; it exists so that the reference's entry
scripts (main.py, train.py, test.py) find every name they import."""
