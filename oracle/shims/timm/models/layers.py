DropPath = to_2tuple = trunc_normal_ = None
