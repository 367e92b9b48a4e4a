function Ws = cnmfe_b200_W2slots(W, tp, tb, r_shift, c_shift, d1, d2)
%% obj.W{m} (sparse d_patch x d_block, initComponents_parallel.m:221-236) -> slot form nnb x d_patch of the library:
% Ws(s, p) = weight of patch pixel p for ring offset (r_shift(s), c_shift(s)); out-of-FOV neighbours stay 0.
% tp / tb: patch_pos / block_pos of the patch (1-based inclusive [r0 r1 c0 c1]); for bg_ssub > 1 pass the coarse grid as both
% ([1 d1s 1 d2s]) with d1 = d1s, d2 = d2s (initComponents_parallel.m:237-253).
nr = tp(2) - tp(1) + 1;  nc = tp(4) - tp(3) + 1;  nrb = tb(2) - tb(1) + 1;
[cc, rr] = meshgrid(tp(3):tp(4), tp(1):tp(2));
rr = rr(:);  cc = cc(:);
nnb = numel(r_shift);
Ws = zeros(nnb, nr * nc);
[ii, jj, vv] = find(W);
Wmap = sparse(ii, jj, vv, size(W, 1), size(W, 2));
for s = 1:nnb
    r2 = rr + double(r_shift(s));  c2 = cc + double(c_shift(s));
    ok = (r2 >= 1) & (r2 <= d1) & (c2 >= 1) & (c2 <= d2);
    jb = (c2 - tb(3)) * nrb + (r2 - tb(1) + 1);
    p = find(ok);
    if isempty(p); continue; end
    idx = sub2ind(size(Wmap), p, jb(ok));
    Ws(s, p) = full(Wmap(idx));
end
end
