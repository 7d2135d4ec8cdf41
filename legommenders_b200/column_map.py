"""Column-name schema (mirror of loader/column_map.py:24-109)."""


class ColumnMap:
    def __init__(self, history_col='history', item_col='item_id', label_col='click', user_col='user_id',
                 group_col='user_id', neg_col=None):
        self.history_col, self.item_col, self.label_col = history_col, item_col, label_col
        self.user_col, self.group_col, self.neg_col = user_col, group_col, neg_col
        self.mask_col = '__clicks_mask__'
        self.history_vocab = self.item_vocab = self.label_vocab = self.user_vocab = self.group_vocab = None

    def set_column_vocab(self, inter_ut):
        feats = inter_ut.meta.features
        self.history_vocab = feats[self.history_col].tokenizer.vocab.name
        self.item_vocab = feats[self.item_col].tokenizer.vocab.name
        self.label_vocab = feats[self.label_col].tokenizer.vocab.name
        self.user_vocab = feats[self.user_col].tokenizer.vocab.name
        self.group_vocab = feats[self.group_col].tokenizer.vocab.name
