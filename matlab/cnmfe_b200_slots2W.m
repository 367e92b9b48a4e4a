function W = cnmfe_b200_slots2W(Ws, tp, tb, r_shift, c_shift, d1, d2)
%% slot form nnb x d_patch -> sparse d_patch x d_block, same pattern as the reference's W (every in-FOV ring neighbour,
% initComponents_parallel.m:221-236; fit_ring_model.m:107 adds 1e-100 so that fitted zeros stay in the pattern).
nr = tp(2) - tp(1) + 1;  nc = tp(4) - tp(3) + 1;
nrb = tb(2) - tb(1) + 1; ncb = tb(4) - tb(3) + 1;
[cc, rr] = meshgrid(tp(3):tp(4), tp(1):tp(2));
rr = rr(:);  cc = cc(:);
nnb = numel(r_shift);
ii = cell(nnb, 1);  jj = cell(nnb, 1);  vv = cell(nnb, 1);
for s = 1:nnb
    r2 = rr + double(r_shift(s));  c2 = cc + double(c_shift(s));
    ok = (r2 >= 1) & (r2 <= d1) & (c2 >= 1) & (c2 <= d2);
    jb = (c2 - tb(3)) * nrb + (r2 - tb(1) + 1);
    ii{s} = find(ok);  jj{s} = jb(ok);  vv{s} = reshape(Ws(s, ok), [], 1);
end
W = sparse(cell2mat(ii), cell2mat(jj), cell2mat(vv), nr * nc, nrb * ncb);
end
