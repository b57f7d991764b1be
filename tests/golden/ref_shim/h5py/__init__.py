class File:  # stub: dataset not available offline
    def __init__(self,*a,**k): raise RuntimeError('h5py stub')
