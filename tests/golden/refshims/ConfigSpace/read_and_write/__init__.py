pcs = pcs_new = json = None
