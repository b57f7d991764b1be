from collections import namedtuple
Genotype = namedtuple('Genotype', 'normal normal_concat reduce reduce_concat')
PRIMITIVES_DEEPNETS1M = ['max_pool','avg_pool','sep_conv','dil_conv','conv','msa','cse','sum','concat','input','bias','bn','ln','pos_enc','glob_avg']
def from_dict(g): return Genotype(normal=g['normal'], normal_concat=g['normal_concat'], reduce=g['reduce'], reduce_concat=g['reduce_concat'])
