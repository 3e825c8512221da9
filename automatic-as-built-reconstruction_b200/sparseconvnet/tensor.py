"""SparseConvNetTensor: (features, metadata, spatial_size) triple passed between scn layers
(reference: SparseConvNet/sparseconvnet/sparseConvNetTensor.py:12-55)."""


class SparseConvNetTensor(object):
    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def get_spatial_locations(self, spatial_size=None):
        """CPU Long [nActive, dim+1]: coordinates and sample index, row-aligned with .features"""
        return self.metadata.getSpatialLocations(self.spatial_size if spatial_size is None else spatial_size)

    def to(self, device):
        self.features = self.features.to(device)
        return self

    def type(self, t=None):
        if t:
            self.features = self.features.type(t)
            return self
        return self.features.type()

    def cuda(self):
        self.features = self.features.cuda()
        return self

    def cpu(self):
        self.features = self.features.cpu()
        return self

    @property
    def requires_grad(self):
        return self.features.requires_grad

    def __repr__(self):
        loc = self.get_spatial_locations() if self.metadata else None
        return ("SparseConvNetTensor<<features=%r,features.shape=%r,batch_locations=%r,"
                "batch_locations.shape=%r,spatial size=%r>>" %
                (self.features, self.features.shape, loc, None if loc is None else loc.shape,
                 self.spatial_size))
