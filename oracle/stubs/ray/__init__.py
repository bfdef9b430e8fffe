"""TEST INFRASTRUCTURE -- `ray` stub so /root/reference/mrsgym/MRSWrapper.py imports."""
