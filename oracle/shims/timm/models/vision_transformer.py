def _cfg(**k):
    return k
