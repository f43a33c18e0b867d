import torch

from .utils import to_undirected


class BaseTransform:
    pass


class FaceToEdge(BaseTransform):
    def __init__(self, remove_faces=True):
        self.remove_faces = remove_faces

    def __call__(self, data):
        face = data.face
        ei = torch.cat([face[:2], face[1:], face[::2]], dim=1)
        data.edge_index = to_undirected(ei, num_nodes=data.num_nodes)
        if self.remove_faces:
            data.face = None
        return data
