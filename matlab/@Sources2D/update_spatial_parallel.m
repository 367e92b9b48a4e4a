function update_spatial_parallel(obj, use_parallel, update_sn) %#ok<INUSL>
%% drop-in for ca_source_extraction/@Sources2D/update_spatial_parallel.m on B200 (update_sn = true is not built).
if exist('update_sn', 'var') && ~isempty(update_sn) && update_sn
    error('cnmfe:b200', 'update_sn=true is not available in the B200 path');
end
h = cnmfe_b200_handle(obj);
search_method = obj.options.search_method;
if strcmpi(search_method, 'dilate'); obj.options.se = []; end
IND = sparse(logical(determine_search_location(obj.A, search_method, obj.options)));   % stays MATLAB (:66)
cnmfe_b200_mex('set_neurons', h, obj.A, obj.C);
cnmfe_b200_mex('set_prev', h, obj.A_prev, obj.C_prev);
cnmfe_b200_mex('set_sn', h, obj.P.sn);
cnmfe_b200_mex('set_search', h, IND);
vals = cnmfe_b200_mex('update_spatial', h, nnz(IND));
[ii, jj] = find(IND);
A_new = sparse(ii, jj, vals, size(IND,1), size(IND,2));
obj.A = obj.post_process_spatial(obj.reshape(A_new, 2));                                  % stays MATLAB (:341)
if strcmpi(obj.options.background_model, 'ring')
    obj.b0_new = cell2mat(obj.P.Ymean) - obj.reshape(obj.A*mean(obj.C,2), 2);
end
flog = fopen(obj.P.log_file, 'a');
fprintf(flog, '[%s]\b', get_minute());
fprintf(flog, 'Finished updating spatial components.\n');
fclose(flog);
end
