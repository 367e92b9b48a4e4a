function cnmfe_b200_save_intermediate(obj, flog, prefix, data)
%% options.save_intermediate (update_background_parallel.m:324-333, update_spatial_parallel.m:357-366,
% update_temporal_parallel.m:300-311): append the snapshot to the matfile obj.P.log_data as <prefix>_<date> and note it in the log.
if ~obj.options.save_intermediate; return; end
log_data = matfile(obj.P.log_data, 'Writable', true);
tmp_str = strrep(get_date(), '-', '_');
log_data.(sprintf('%s_%s', prefix, tmp_str)) = data;
fprintf(flog, '\tThe results were saved as intermediate_results.%s_%s\n\n', prefix, tmp_str);
end
