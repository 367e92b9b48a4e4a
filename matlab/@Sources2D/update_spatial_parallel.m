function update_spatial_parallel(obj, use_parallel, update_sn) %#ok<INUSL>
%% drop-in for ca_source_extraction/@Sources2D/update_spatial_parallel.m on B200.
if ~exist('update_sn', 'var') || isempty(update_sn); update_sn = false; end
h = cnmfe_b200_push(obj, {'neurons', 'prev', 'sn'});   % options + background state every time: they may have changed since the last call
search_method = obj.options.search_method;
if strcmpi(search_method, 'dilate'); obj.options.se = []; end
% determine_search_location(obj.A, search_method, options) (:66) in the library (host C++, same results)
if strcmpi(search_method, 'ellipse') && ~isinf(obj.options.dist)
    [jc, ir] = cnmfe_b200_mex('search_location', obj.A, obj.options.d1, obj.options.d2, obj.options.min_size, obj.options.max_size, obj.options.dist);
elseif strcmpi(search_method, 'dilate')
    [jc, ir] = cnmfe_b200_mex('search_location_dilate', obj.A, obj.options.d1, obj.options.d2, obj.options.nrgthr, obj.options.nb, obj.options.bSiz);
else
    jc = [];
end
if ~isempty(jc)
    jj = zeros(numel(ir), 1); for k = 1:numel(jc)-1; jj(jc(k)+1:jc(k+1)) = k; end
    IND = sparse(ir + 1, jj, true, size(obj.A,1), size(obj.A,2));
else
    IND = sparse(logical(determine_search_location(obj.A, search_method, obj.options)));   % whole field of view
end
cnmfe_b200_mex('set_search', h, IND);
if update_sn
    [vals, obj.P.sn] = cnmfe_b200_mex('update_spatial', h, nnz(IND), true, obj.options.d1, obj.options.d2);   % (:191-194, :337-339)
else
    vals = cnmfe_b200_mex('update_spatial', h, nnz(IND));
end
[ii, jj] = find(IND);
A_new = sparse(ii, jj, vals, size(IND,1), size(IND,2));
sc = obj.options.spatial_constraints;                                                          % post_process_spatial (:341) in the library
if sc.connected
    Atmp = cnmfe_b200_mex('post_process_spatial', A_new, obj.options.d1, obj.options.d2);      % connectivity_constraint per neuron
    [ii2, jj2, vv2] = find(Atmp);                                                               % removed entries come back as explicit zeros
    A_new = sparse(ii2, jj2, vv2, size(Atmp,1), size(Atmp,2));
end
if sc.circular
    [jc, ir, vv] = cnmfe_b200_mex('circular_constraints', A_new, obj.options.d1, obj.options.d2);   % circular_constraints per neuron
    jj = zeros(numel(ir), 1); for k = 1:numel(jc)-1; jj(jc(k)+1:jc(k+1)) = k; end
    A_new = sparse(ir + 1, jj, vv, size(A_new,1), size(A_new,2));
end
obj.A = A_new;
if strcmpi(obj.options.background_model, 'ring')
    obj.b0_new = cell2mat(obj.P.Ymean) - obj.reshape(obj.A*mean(obj.C,2), 2);
end
flog = fopen(obj.P.log_file, 'a');
fprintf(flog, '[%s]\b', get_minute());
fprintf(flog, 'Finished updating spatial components.\n');
spatial = struct('A', obj.A, 'ids', obj.ids);
cnmfe_b200_save_intermediate(obj, flog, 'spatial', spatial);
fclose(flog);
end
