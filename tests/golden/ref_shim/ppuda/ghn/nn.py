import copy, numpy as np, torch, torch.nn as nn
from ..deepnets1m.genotypes import PRIMITIVES_DEEPNETS1M
def get_activation(a): return nn.ReLU() if a == 'relu' else nn.Identity()
class MLP(nn.Module):
    def __init__(self, in_features=32, hid=(32,32), activation='relu', last_activation='same'):
        super().__init__(); fc=[]; n_hid=len(hid)
        for j,n in enumerate(hid):
            fc += [nn.Linear(in_features if j==0 else hid[j-1], n),
                   get_activation(last_activation if (j==n_hid-1 and last_activation!='same') else activation)]
        self.fc = nn.Sequential(*fc)
    def forward(self, x, *a, **k): return self.fc(x[0] if isinstance(x, tuple) else x)
class ConvDecoder(nn.Module):
    def __init__(self, in_features=64, hid=(128,256), out_shape=None, num_classes=None):
        super().__init__(); self.out_shape=out_shape; self.num_classes=num_classes
        self.fc = nn.Sequential(nn.Linear(in_features, hid[0]*int(np.prod(out_shape[2:]))), nn.ReLU())
        conv=[]
        for j,n_hid in enumerate(hid):
            n_out = int(np.prod(out_shape[:2])) if j==len(hid)-1 else hid[j+1]
            conv += [nn.Conv2d(n_hid, n_out, 1), get_activation(None if j==len(hid)-1 else 'relu')]
        self.conv = nn.Sequential(*conv)
        self.class_layer_predictor = nn.Sequential(nn.ReLU(), nn.Conv2d(out_shape[0], num_classes, 1))
class ShapeEncoder(nn.Module):
    def __init__(self, hid, num_classes, max_shape, debug_level=0):
        super().__init__(); self.debug_level=debug_level; self.num_classes=num_classes
        self.ch_steps=(2**3,2**6,2**12,2**13)
        self.channels=np.unique([1,3,num_classes]+list(range(8,64,8))+list(range(64,4096,16))+list(range(4096,8193,32)))
        self.spatial=np.unique(list(range(1,max(12,max_shape[3]),2))+[14,16])
        self.channels_lookup={c:i for i,c in enumerate(self.channels)}
        for c in range(4,8): self.channels_lookup[c]=self.channels_lookup[8]
        for c in range(1,self.channels[-1]):
            if c not in self.channels_lookup:
                self.channels_lookup[c]=self.channels_lookup[self.channels[np.argmin(abs(self.channels-c))]]
        self.spatial_lookup={c:i for i,c in enumerate(self.spatial)}
        self.spatial_lookup[2]=self.spatial_lookup[3]
        for c in range(1,self.spatial[-1]):
            if c not in self.spatial_lookup:
                self.spatial_lookup[c]=self.spatial_lookup[self.spatial[np.argmin(abs(self.spatial-c))]]
        n_ch,n_s=len(self.channels),len(self.spatial)
        self.embed_spatial=nn.Embedding(n_s+1,hid//4); self.embed_channel=nn.Embedding(n_ch+1,hid//4)
        self.register_buffer('dummy_ind', torch.tensor([n_ch,n_ch,n_s,n_s],dtype=torch.long).view(1,4), persistent=False)
    def forward(self, x, params_map, predict_class_layers=True):
        shape_ind=self.dummy_ind.repeat(len(x),1)
        for node_ind in params_map:
            sz=params_map[node_ind][0]['sz']
            if sz is None: continue
            if len(sz)==1: sz=(sz[0],1)
            if len(sz)==2: sz=(sz[0],sz[1],1,1)
            if len(sz)==3: sz=(sz[0],sz[1],sz[2],1)   # NOTE: guess for 3d (layer_scale); real ppuda behaviour unknown
            for i in range(4):
                if i<2: shape_ind[node_ind,i]=self.channels_lookup[sz[i] if sz[i] in self.channels_lookup else self.channels[-1]]
                else: shape_ind[node_ind,i]=self.spatial_lookup[sz[i] if sz[i] in self.spatial_lookup else self.spatial[-1]]
        e=torch.cat((self.embed_channel(shape_ind[:,0]),self.embed_channel(shape_ind[:,1]),
                     self.embed_spatial(shape_ind[:,2]),self.embed_spatial(shape_ind[:,3])),dim=1)
        return x+e
class GHN(nn.Module):
    def __init__(self, max_shape, num_classes, hypernet='gatedgnn', decoder='conv', weight_norm=False, ve=False,
                 layernorm=False, hid=32, debug_level=0):
        super().__init__()
        self.max_shape=max_shape; self.layernorm=layernorm; self.weight_norm=weight_norm; self.ve=ve
        self.debug_level=debug_level; self.num_classes=num_classes
        if layernorm: self.ln=nn.LayerNorm(hid)
        self.embed=nn.Embedding(len(PRIMITIVES_DEEPNETS1M),hid)
        self.shape_enc=ShapeEncoder(hid=hid,num_classes=num_classes,max_shape=max_shape,debug_level=debug_level)
        self.gnn=nn.Identity()
        self.decoder=ConvDecoder(in_features=hid,hid=(hid*4,hid*8),out_shape=max_shape,num_classes=num_classes)
        max_ch=max(max_shape[:2])
        self.decoder_1d=MLP(hid,hid=(hid*2,2*max_ch),last_activation=None)
        self.bias_class=nn.Sequential(nn.ReLU(),nn.Linear(max_ch,num_classes))
